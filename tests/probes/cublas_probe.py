"""What does the vendor GEMM (cuBLAS through torch.matmul) achieve on the shapes of this path?  Reference points for
the roofline claims only -- nothing in the product calls cuBLAS.

    python tests/probes/cublas_probe.py time      # CUDA-event timings, 3 + 10 calls per shape
    python tests/probes/cublas_probe.py once      # one call per shape (run under ncu)
"""
import sys

import torch

SHAPES = [
    # name, (m, k) A, (k, n) B as "nt" (B given as (n, k), transposed view) or "nn"
    ("peak_8192^3_bf16_nt", 8192, 8192, 8192, torch.bfloat16, "nt"),
    ("similarity_32768x32768x768_f16_nt", 32768, 32768, 768, torch.float16, "nt"),
    ("grad_rows_32768x768x32768_f16_nn", 32768, 768, 32768, torch.float16, "nn"),
    ("grad_cols_32768x768x32768_f16_tn", 32768, 768, 32768, torch.float16, "tn"),
]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "time"
    dev = "cuda"
    for name, m, n, k, dt, lay in SHAPES:
        if lay == "nt":
            a = torch.randn(m, k, device=dev, dtype=dt)
            b = torch.randn(n, k, device=dev, dtype=dt).t()
        elif lay == "nn":
            a = torch.randn(m, k, device=dev, dtype=dt)
            b = torch.randn(k, n, device=dev, dtype=dt)
        else:
            a = torch.randn(k, m, device=dev, dtype=dt).t()
            b = torch.randn(k, n, device=dev, dtype=dt)
        out = torch.empty(m, n, device=dev, dtype=dt)
        if mode == "once":
            torch.matmul(a, b, out=out)
            torch.cuda.synchronize()
            continue
        for _ in range(3):
            torch.matmul(a, b, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch.matmul(a, b, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"CUBLAS {name}: {ms:.3f} ms  {2.0 * m * n * k / (ms * 1e-3) / 1e12:.1f} TFLOP/s", flush=True)
        del a, b, out


if __name__ == "__main__":
    main()
