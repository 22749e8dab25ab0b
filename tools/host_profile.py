import os, sys, cProfile, pstats, io, time
sys.path.insert(0, os.getcwd())
import torch, vpr_b200
B, n, hw = 256, 256, 16
torch.manual_seed(0)
head = vpr_b200.SaladHead(dropout=0.3).cuda().train()
x = torch.randn(B, n, 768, device="cuda").bfloat16()
t = torch.randn(B, 768, device="cuda").bfloat16()
w = torch.randn(B, 8448, device="cuda").bfloat16()
def step():
    for p in head.parameters():
        p.grad = None
    xg = x.detach().requires_grad_(True)
    tg = t.detach().requires_grad_(True)
    out = head((xg.view(B, hw, hw, 768).permute(0, 3, 1, 2), tg))
    out.backward(w)
for _ in range(5): step()
torch.cuda.synchronize()
# forward-only and backward-only host timings
t0 = time.perf_counter()
outs = []
for _ in range(30):
    xg = x.detach().requires_grad_(True); tg = t.detach().requires_grad_(True)
    outs.append(head((xg.view(B, hw, hw, 768).permute(0, 3, 1, 2), tg)))
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
for o in outs: o.backward(w)
t3 = time.perf_counter()
torch.cuda.synchronize()
print(f"host: forward {1e6*(t1-t0)/30:.0f} us, backward {1e6*(t3-t2)/30:.0f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(100): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:5000])
