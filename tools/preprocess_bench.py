"""GPU preprocessing throughput: a batch of equal-sized uint8 RGB images -> [B,3,224,224] (Pillow-exact resize + normalise).
    python tools/preprocess_bench.py [B] [H] [W]
Reports images/s and the bytes the two passes move (input + intermediate write/read + output) against the HBM peak,
plus the single-thread Pillow + torchvision time for the same transform on the host for reference."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vpr_b200  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 480
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 640
    imgs = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda")
    for dt in (torch.float32, torch.bfloat16):
        pre = vpr_b200.GpuPreprocess((224, 224), out_dtype=dt)
        for _ in range(3):
            pre(imgs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = pre(imgs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        byt = B * (H * W * 3 + 2 * H * 224 * 3 + 3 * 224 * 224 * out.element_size())
        print(json.dumps({"batch": B, "in": [H, W], "out_dtype": str(dt), "ms": ms, "img_per_s": B / ms * 1e3, "GB/s": byt / ms / 1e6}))
    try:
        from PIL import Image
        from torchvision import transforms
        tf = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(),
                                 transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        pil = [Image.fromarray(imgs[i].cpu().numpy(), "RGB") for i in range(min(B, 32))]
        t0 = time.time()
        for im in pil:
            tf(im)
        dt_ = time.time() - t0
        print(json.dumps({"host torchvision+Pillow (the reference's transform, one thread)": len(pil) / dt_, "unit": "img/s"}))
    except Exception as e:   # noqa: BLE001
        print("host reference not available:", e)


if __name__ == "__main__":
    main()
