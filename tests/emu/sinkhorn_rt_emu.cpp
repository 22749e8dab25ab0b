// C entry points that run csrc/sinkhorn_rt.cuh on the CPU through tests/emu/simt_emu.h (test infrastructure only).
#include "simt_emu.h"

#include "../../visual-place-recognition-and-geopose-estimation_b200/csrc/sinkhorn_rt.cuh"

namespace {
vpr::SinkConst2 consts(int n) {
    vpr::SinkConst2 k;
    k.norm2 = (float)(-std::log2((double)(n + 64)));
    k.log_a_m2 = k.norm2 + (float)std::log2((double)(n - 64));
    return k;
}
}  // namespace

extern "C" {
// returns the number of images that asked for the log-domain guard path (their outputs are then not written)
// nt: threads per block, 512 (two per token row) or 256 (one per row)
int emu_sinkhorn_fwd_rt(const float* S, const float* dust_bin, void* P, int p_bf16, float* U, float* V, int B, int n, int iters, int nt) {
    simt_emu::guard_count = 0;
    const vpr::SinkConst2 k = consts(n);
    const size_t smem = vpr::sinkrt_fwd_floats() * sizeof(float);
    if (nt != 512 && nt != 256) return -1;
    if (p_bf16 && nt == 512)
        simt_emu::launch(B, nt, smem, [=]() { vpr::sinkhorn_fwd_rt<__nv_bfloat16, 512>(S, dust_bin, (__nv_bfloat16*)P, U, V, n, iters, k); });
    else if (p_bf16)
        simt_emu::launch(B, nt, smem, [=]() { vpr::sinkhorn_fwd_rt<__nv_bfloat16, 256>(S, dust_bin, (__nv_bfloat16*)P, U, V, n, iters, k); });
    else if (nt == 512)
        simt_emu::launch(B, nt, smem, [=]() { vpr::sinkhorn_fwd_rt<float, 512>(S, dust_bin, (float*)P, U, V, n, iters, k); });
    else
        simt_emu::launch(B, nt, smem, [=]() { vpr::sinkhorn_fwd_rt<float, 256>(S, dust_bin, (float*)P, U, V, n, iters, k); });
    return simt_emu::guard_count.load();
}
int emu_sinkhorn_bwd_rt3(const float* S, const float* dust_bin, const float* U, const float* V, const float* dP, void* dS, int ds_bf16,
                         float* d_bias, int B, int n, int nt) {
    simt_emu::guard_count = 0;
    const vpr::SinkConst2 k = consts(n);
    const size_t smem = vpr::sinkrt_bwd_floats(3) * sizeof(float);
    if (nt != 512 && nt != 256) return -1;
    if (ds_bf16 && nt == 512)
        simt_emu::launch(B, nt, smem, [=]() { vpr::sinkhorn_bwd_rt<__nv_bfloat16, 3, 512>(S, dust_bin, U, V, dP, (__nv_bfloat16*)dS, d_bias, n, k); });
    else if (ds_bf16)
        simt_emu::launch(B, nt, smem, [=]() { vpr::sinkhorn_bwd_rt<__nv_bfloat16, 3, 256>(S, dust_bin, U, V, dP, (__nv_bfloat16*)dS, d_bias, n, k); });
    else if (nt == 512)
        simt_emu::launch(B, nt, smem, [=]() { vpr::sinkhorn_bwd_rt<float, 3, 512>(S, dust_bin, U, V, dP, (float*)dS, d_bias, n, k); });
    else
        simt_emu::launch(B, nt, smem, [=]() { vpr::sinkhorn_bwd_rt<float, 3, 256>(S, dust_bin, U, V, dP, (float*)dS, d_bias, n, k); });
    return simt_emu::guard_count.load();
}
// Forward with the plan ALSO written as the aggregation MMA's shared-memory B operand (what the fused Sinkhorn +
// aggregation kernel does); the operand bytes of every image ([op_rows / 128] chunks of 16 KB) are copied out to `op_out`.
int emu_sinkhorn_fwd_rt_operand(const float* S, const float* dust_bin, void* P, float* U, float* V, unsigned char* op_out, int B, int n,
                                int iters) {
    simt_emu::guard_count = 0;
    const vpr::SinkConst2 k = consts(n);
    const size_t smem = vpr::sinkrt_fwd_floats() * sizeof(float);
    const int op_rows = (n + 127) / 128 * 128;
    simt_emu::launch(B, 512, smem, [=]() {
        float* sk = simt_emu::dyn_smem();
        vpr::RtFwdState<512> st;
        if (vpr::rt_fwd_iterate<512>(sk, S, dust_bin, U, V, n, iters, k, st)) { simt_emu::note_guard(); return; }
        uint8_t* op = reinterpret_cast<uint8_t*>(sk);            // over the dead tile, as in the fused kernel
        vpr::rt_fwd_plan<__nv_bfloat16, 512, true>(sk + vpr::RT_ROWS * vpr::RT_PITCH + 2 * vpr::RT_ROWS, st, (__nv_bfloat16*)P, n, op, op_rows);
        __syncthreads();
        for (int i = threadIdx.x; i < op_rows * 128; i += 512) op_out[(size_t)blockIdx.x * op_rows * 128 + i] = op[i];
    });
    return simt_emu::guard_count.load();
}
size_t emu_sinkrt_fwd_bytes() { return vpr::sinkrt_fwd_floats() * sizeof(float); }
size_t emu_sinkrt_bwd_bytes(int iters) { return vpr::sinkrt_bwd_floats(iters) * sizeof(float); }
}
