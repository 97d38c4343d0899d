#!/usr/bin/env python
"""Benchmark of the fused tri-modal contrastive objective (BASELINE.json metric: fwd+bwd samples/s, % of
tensor-core peak).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, north-star workload
    python bench.py --workload cfg1|cfg2|north_star          # BASELINE.json configs 1-3 as the headline of the line
    python bench.py --workload sweep [--gpus N]              # BASELINE config 5: B x D sweep, one JSON line per point
                                                             # into gpurun_out/sweep_w<N>.jsonl (+ one summary line)
    python bench.py --workload step [--gpus N]               # BASELINE config 4: full Base pre-training step (encoders +
                                                             # loss tail, DDP, AdamW) on synthetic inputs, six arms
    python bench.py --impl reference [--workload ...]        # the reference's PyTorch CPU loss path (oracle port)

One "step" = one forward + backward of the loss tail (model.py:247-272 + autograd) over one synthetic global
batch.  Default workload (every N, strong scaling): the north-star shape B=32768, D=768, bf16 embeddings, row-sharded
over the N ranks; at N=1 the same run also reports config 2 (8192 x 512 bf16), config 1 (256 x 512 fp32, with the CPU
baseline timed for real on that shape: 5 + 30 iterations), the CPU baseline of the north-star
shape (a bounded sub-batch, EXTRAPOLATED, labelled as such) and the reference's own tail under torch eager on the same
GPU (`gpu_eager_baseline`: "the existing kernels on the same box").  Prints exactly one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOGIT_SCALE_INIT = 2.6592  # config.py:112
WORKLOADS = {"north_star": (32768, 768, "bf16"), "cfg2": (8192, 512, "bf16"), "cfg1": (256, 512, "f32")}
CPU_SAMPLE_ROWS = 4096  # bounded CPU sample of the large workloads: a 4096-row sub-batch of the same embeddings
SWEEP_B = (4096, 8192, 16384, 32768, 65536, 131072)
SWEEP_D = (512, 768, 1024)


def load_traffic(b, d, world):
    """dram__bytes_read + dram__bytes_write of the dominant kernel from the committed `ncu --set full` capture (a
    profiler figure cannot be taken inside a timed run; the capture is of this workload at one GPU only)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if t.get("rows_global") == b and t.get("dim") == d and t.get("world") == world:
            return t.get("gemm_wide_kernel_dram_bytes")
    except Exception:  # noqa: BLE001
        pass
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU (NVML, ~20 ms period; nvidia-smi as a fallback)
    while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []  # (sm_mhz, sm_max_mhz, power_w, reasons bitmask)
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self):
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        while not self._stop.is_set():
            self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), float(mx),
                                 pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                 int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
            self._stop.wait(0.02)

    def _run_smi(self):
        query = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
        bits = (0x8, 0x40, 0x20, 0x4)
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={query}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    mask = sum(b for b, v in zip(bits, parts[3:7]) if v.lower() == "active")
                    self.samples.append((float(parts[0]), float(parts[1]), float(parts[2]), mask))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def _run(self):
        try:
            self._run_nvml()
        except Exception:  # noqa: BLE001
            self._run_smi()

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        mask = 0
        for s in self.samples:
            mask |= s[3]
        names = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                 (0x4, "sw_power_cap"))
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1],
                "reasons": [n for b, n in names if mask & b], "samples": len(sm),
                "power_w_max": max(s[2] for s in self.samples)}


def cpu_tail_rate(rows: int, dim: int, steps: int, warmup: int, seed: int = 1234):
    """The reference's PyTorch CPU loss path (oracle/reference_tail.py restates model.py:247-272 verbatim) in fp32
    on all host cores: returns (samples/s, seconds per step, threads)."""
    import torch

    from oracle import reference_tail

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    leaves = [torch.randn(rows, dim, generator=g).requires_grad_(True) for _ in range(3)]
    scales = [torch.tensor(LOGIT_SCALE_INIT, requires_grad=True) for _ in range(3)]
    for _ in range(warmup):
        reference_tail.timed_step(*leaves, scales)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_tail.timed_step(*leaves, scales)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return rows / dt, dt, threads


def cpu_full_batch_rate(rows: int, b: int, dt: float) -> float:
    """The work of the tail is quadratic in the batch (3 BxB similarity matrices): a step over the full batch of b
    samples costs (b / rows)^2 sub-batch steps, so the whole-job rate on the full workload is b / (dt (b/rows)^2)."""
    return b / (dt * (b / rows) ** 2)


def cpu_block(b: int, d: int, steps: int, warmup: int):
    """cpu_baseline object for a workload.  b <= CPU_SAMPLE_ROWS: the reference functions timed for real on that shape.
    Larger: a CPU_SAMPLE_ROWS-row sub-batch (the full batch needs nine B x B fp32 matrices: 36 GiB at 32768), scaled
    by the quadratic work ratio and LABELLED extrapolated."""
    rows = min(CPU_SAMPLE_ROWS, b)
    sub_rate, dt, threads = cpu_tail_rate(rows, d, steps, warmup)
    if rows == b:
        return {"value": sub_rate, "unit": "samples/s", "cores": threads, "kind": "port", "extrapolated": False,
                "ms_per_step": dt * 1e3,
                "sample": f"the whole {b}x{d} workload, fp32 torch CPU on {threads} threads, {warmup} warm-up + {steps} "
                          f"timed fwd+bwd steps of the reference's statements (oracle/reference_tail.py = "
                          f"model.py:247-272), {dt * 1e3:.2f} ms/step"}, dt
    rate = cpu_full_batch_rate(rows, b, dt)
    return {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port", "extrapolated": True,
            "ms_per_step": dt * 1e3 * (b / rows) ** 2,
            "sample": f"EXTRAPOLATED: {warmup} warm-up + {steps} timed fwd+bwd steps of a {rows}-row sub-batch of the "
                      f"{b}x{d} workload (fp32 torch CPU, {threads} threads, {dt:.2f} s/step = {sub_rate:.0f} samples/s "
                      f"at B={rows}), scaled by the quadratic work ratio (B/{rows})^2; the full batch itself needs nine "
                      f"B x B fp32 matrices"}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload if args.workload in WORKLOADS else "north_star"
    b, d, _ = WORKLOADS[name]
    if b <= CPU_SAMPLE_ROWS:  # config 1: 5 warm-up + 30 timed iterations of the real shape
        steps, warmup = max(args.steps, 30), max(args.warmup, 5)
    else:
        steps, warmup = args.steps, args.warmup
    cpu, dt = cpu_block(b, d, steps, warmup)
    line = {
        "impl": "reference", "metric": "contrastive_loss_fwd_bwd_samples_per_sec", "value": cpu["value"],
        "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": cpu["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"tri-modal contrastive loss fwd+bwd, global batch {b}, dim {d}", "rows_global": b,
                   "dim": d, "reference_sample_rows": min(CPU_SAMPLE_ROWS, b), "extrapolated": cpu["extrapolated"]},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def time_steps(fn, steps, warmup, dist_mod=None):
    """W warm-up + K timed calls of fn, CUDA events on the current stream, barrier + synchronize on both sides,
    max over ranks.  Returns ms per step."""
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist_mod is not None:
        dist_mod.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if dist_mod is not None:
        dist_mod.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist_mod is not None:
        dist_mod.all_reduce(ms, op=dist_mod.ReduceOp.MAX)
    return ms.item() / steps


def stage_breakdown(run, steps):
    """Average device time of every stage (CUDA events recorded between the stage launches of `run`)."""
    import torch

    from synergy_clip_b200 import ops

    acc = {}
    for _ in range(steps):
        marks = []

        def trace(name, marks=marks):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

        ops._TRACE = trace
        try:
            run()
        finally:
            ops._TRACE = None
        torch.cuda.synchronize()
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            if name in ("begin", "backward_begin"):
                continue
            acc.setdefault(name, []).append(a.elapsed_time(b))
    avg = {k: sum(v) / len(v) for k, v in acc.items()}
    # the sharded path marks sub-stages (the two GEMM roles, the statistics exchange): fold them into the six stage
    # names, keeping the detail under "<stage>/<sub>"
    folds = {"forward_reduce": "forward_finish", "forward_barrier1": "forward_finish",
             "backward_gemms_col": "backward_gemms", "backward_gemms_row": "backward_gemms"}
    out = {}
    for k, v in avg.items():
        tgt = folds.get(k, k)
        out[tgt] = out.get(tgt, 0.0) + v
        if tgt != k:
            out[f"{tgt}/{k}"] = v
    return out


def reference_tail_eager(img, txt, aud, scales):
    """The reference's statements (model.py:247-272 with clip_loss / contrastive_loss, model.py:52-58) on whatever
    device and dtype the tensors have -- what the unmodified reference runs under torch eager, fwd + bwd."""
    import torch
    import torch.nn.functional as F

    for p in (img, txt, aud, *scales):
        p.grad = None
    img_n = img / img.norm(p=2, dim=-1, keepdim=True)
    txt_n = txt / txt.norm(p=2, dim=-1, keepdim=True)
    aud_n = aud / aud.norm(p=2, dim=-1, keepdim=True)
    total = 0
    for a, b, t in ((img_n, txt_n, scales[0]), (txt_n, aud_n, scales[1]), (aud_n, img_n, scales[2])):
        sim = torch.matmul(a, b.t()) * t.exp()
        labels = torch.arange(sim.shape[0], device=sim.device)
        total = total + (F.cross_entropy(sim, labels) + F.cross_entropy(sim.t(), labels)) / 2.0
    total.backward()
    return total


def gpu_eager_block(dev, shapes, ours_ms):
    """`gpu_eager_baseline`: the reference's own tail under torch eager (cuBLAS + ATen kernels) on this GPU, fp32 (what
    the reference runs) and bf16 (what it would run under autocast-style casting), CUDA events, 1 warm-up + 3 timed."""
    import torch

    out = {}
    for name, (b, d) in shapes.items():
        g = torch.Generator(device=dev).manual_seed(1234)
        base = [torch.randn(b, d, device=dev, generator=g) for _ in range(3)]
        entry = {"rows": b, "dim": d, "fused_ms": ours_ms.get(name)}
        for label, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
            try:
                leaves = [x.to(dt).requires_grad_(True) for x in base]
                scales = [torch.tensor(LOGIT_SCALE_INIT, device=dev, requires_grad=True) for _ in range(3)]
                torch.cuda.reset_peak_memory_stats(dev)
                reference_tail_eager(*leaves, scales)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    reference_tail_eager(*leaves, scales)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                entry[label] = {"ms_per_step": ms, "samples_per_s": b / (ms * 1e-3),
                                "peak_gib": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                                "fused_speedup": (ms / ours_ms[name]) if ours_ms.get(name) else None}
                del leaves, scales
            except RuntimeError as e:  # out of memory at this shape
                entry[label] = {"error": str(e).split("\n")[0][:160]}
            torch.cuda.empty_cache()
        out[name] = entry
        del base
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="north_star", choices=sorted(WORKLOADS) + ["sweep", "step"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip configs 1 / 2 and the GPU-eager baseline")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: run the collectives on the compute stream")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from synergy_clip_b200 import fused_tri_contrastive, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run for N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    pg = None
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: one JSON line only
        # the collectives run next to persistent tile kernels that leave them 20 SMs (TriContrastiveConfig.comm_sms)
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    peaks = load_peaks()

    def make_inputs(b, d, dtype, shard):
        g = torch.Generator(device=dev).manual_seed(1234)
        full = [torch.randn(b, d, device=dev, generator=g).to(dtype) for _ in range(3)]
        if shard and world > 1:
            bl = b // world
            full = [f[rank * bl:(rank + 1) * bl].contiguous() for f in full]
        return full

    def measure(b, d, dt_name, shard, steps, warmup, with_e2e=True, with_stages=True):
        dtype = torch.float32 if dt_name == "f32" else torch.bfloat16
        cfg = ops.TriContrastiveConfig(process_group=pg if shard else None, math="auto", grad_scale="ddp",
                                       overlap=not args.no_overlap)
        embs = make_inputs(b, d, dtype, shard)
        t3 = torch.full((3,), LOGIT_SCALE_INIT, device=dev)
        g3 = torch.ones(3, device=dev)
        out = {}

        def resident_step():
            out["r"] = ops.forward_backward_raw(*embs, t3, g3, cfg)

        from synergy_clip_b200 import _lib as sclip_lib

        counter = sclip_lib.load().sclip_kernel_launches
        # warm up outside the counted / clock-sampled window, then time exactly `steps` steps
        for _ in range(warmup):
            resident_step()
        torch.cuda.synchronize()
        launches0 = counter()
        sampler = ClockSampler(local_rank)
        with sampler:
            ms = time_steps(resident_step, steps, 0, dist if shard else None)
        launches = counter() - launches0
        stages = stage_breakdown(resident_step, min(steps, 5)) if with_stages else {}
        res = {"ms": ms, "stages": stages, "clocks": sampler.summary(), "launches": launches,
               "loss": [float(x) for x in out["r"][0].tolist()], "rows_local": embs[0].shape[0]}
        if not with_e2e:
            ops._POOL.clear()
            return res

        # end to end through the public autograd API with HOST buffers: every step copies its three embedding matrices
        # from pinned host memory and reads the three losses back.  Like a training loop with a prefetching loader, the
        # copy of step n+1 is issued on a second stream while step n computes (two device buffer sets).
        host = [e.cpu().pin_memory() for e in embs]
        dbuf = [[torch.empty_like(e) for e in embs] for _ in range(2)]
        params = [torch.full((), LOGIT_SCALE_INIT, device=dev, requires_grad=True) for _ in range(3)]
        loss_host = torch.empty(3, dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev)
        arrived = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"cur": 0}

        def prefetch(slot):
            with torch.cuda.stream(copy_stream):
                for h, dst in zip(host, dbuf[slot]):
                    dst.copy_(h, non_blocking=True)
                arrived[slot].record(copy_stream)

        prefetch(0)

        def e2e_step():
            slot = state["cur"]
            torch.cuda.current_stream().wait_event(arrived[slot])
            prefetch(1 - slot)  # its previous consumer finished before the host sync at the end of the last step
            leaves = [d.detach().requires_grad_(True) for d in dbuf[slot]]
            for p in params:
                p.grad = None
            it, ta, ai = fused_tri_contrastive(*leaves, *params, config=cfg)
            (it + ta + ai).backward()
            loss_host.copy_(torch.stack([it.detach(), ta.detach(), ai.detach()]), non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the caller reads the losses (main_pretraining.py:169-170)
            state["cur"] = 1 - slot

        res["ms_e2e"] = time_steps(e2e_step, steps, warmup, dist if shard else None)
        res["h2d"] = sum(h.numel() * h.element_size() for h in host)
        if shard and dist is not None:
            lt = torch.tensor([launches], device=dev, dtype=torch.int64)
            dist.all_reduce(lt)
            res["launches"] = int(lt.item())
        ops._POOL.clear()
        return res

    if args.workload == "sweep":
        return run_sweep(args, measure, world, rank, peaks, dist)
    if args.workload == "step":
        return run_step(args, world, rank, dev, dist)

    b, d, dt_name = WORKLOADS[args.workload]
    shard = args.workload == "north_star"
    if not shard and world > 1:
        raise SystemExit("configs 1 and 2 are single-GPU workloads (BASELINE.json): run them with --gpus 1")
    # The small BASELINE configs are measured FIRST, from an idle GPU: after a long run at the 1 kW power cap the SM
    # clock stays low for a while (a sweep point measured right after a 128k batch ran at 832 MHz), which says nothing
    # about a 1 ms workload.  Their clocks are recorded next to the numbers.
    also_res = {}
    if not args.no_also and rank == 0 and world == 1:
        for name in ("cfg1", "cfg2"):
            if name != args.workload:
                b2, d2, dt2 = WORKLOADS[name]
                also_res[name] = measure(b2, d2, dt2, False, max(args.steps, 20), args.warmup)
    main_res = measure(b, d, dt_name, shard, args.steps, args.warmup)
    flops = 18.0 * b * b * d
    value = b / (main_res["ms"] * 1e-3)
    tflops_per_gpu = flops / world / (main_res["ms"] * 1e-3) / 1e12
    st = main_res["stages"]
    # dominant kernel: the gradient GEMM launch (6 of the 9 contractions = 12 B^2 D / world flops per launch)
    gemm_ms = st.get("backward_gemms")
    gemm_tf = 12.0 * b * b * d / world / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None
    stashed = dt_name == "bf16" and d >= 512
    stage_rates = {}
    if st.get("forward_tiles"):
        stage_rates["forward_tiles_tflops"] = 6.0 * b * b * d / world / (st["forward_tiles"] * 1e-3) / 1e12
    if gemm_tf:
        stage_rates["backward_gemms_tflops"] = gemm_tf
    if st.get("backward_tiles"):
        if stashed:  # the in-place stash -> G' conversion executes no algorithmic flop: an HBM pass, 4 bytes per logit
            gbs = 3.0 * 4.0 * b * b / world / (st["backward_tiles"] * 1e-3) / 1e9
            stage_rates["backward_scale_gbs"] = gbs
            stage_rates["backward_scale_frac_of_hbm_peak"] = gbs / peaks["hbm_gbs"]
        else:        # the recompute backward executes 6 B^2 D flop that the algorithmic count does not include
            stage_rates["backward_tiles_executed_tflops"] = 6.0 * b * b * d / world / (st["backward_tiles"] * 1e-3) / 1e12

    also, eager = None, None
    if not args.no_also and rank == 0 and world == 1:
        also = {}
        ours_ms = {args.workload: main_res["ms"]}
        for name in ("cfg2", "cfg1"):
            if name == args.workload:
                continue
            b2, d2, dt2 = WORKLOADS[name]
            r2 = also_res[name]
            ours_ms[name] = r2["ms"]
            tf2 = 18.0 * b2 * b2 * d2 / (r2["ms"] * 1e-3) / 1e12
            also[f"{name}_{b2}x{d2}_{dt2}_1gpu"] = {
                "value": b2 / (r2["ms"] * 1e-3), "unit": "samples/s", "ms_per_step": r2["ms"],
                "tflops_algorithmic": tf2, "frac_of_bf16_peak": tf2 / peaks["bf16_tflops"],
                "e2e_value": b2 / (r2["ms_e2e"] * 1e-3), "gpu_launches_per_step": r2["launches"] / max(args.steps, 20),
                "clocks": r2["clocks"], "stages_ms": r2["stages"]}
        shapes = {k: WORKLOADS[k][:2] for k in ours_ms}
        eager = gpu_eager_block(dev, shapes, ours_ms)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if b <= CPU_SAMPLE_ROWS:
            cpu, _ = cpu_block(b, d, 30, 5)
        else:
            cpu, _ = cpu_block(b, d, 3, 1)
            if not args.no_also:  # config 1: the reference functions timed for real (5 + 30)
                c1, _ = cpu_block(*WORKLOADS["cfg1"][:2], 30, 5)
                cpu["cfg1_256x512_f32"] = c1

    if rank == 0:
        tstep = load_traffic(b, d, world)
        line = {
            "metric": "contrastive_loss_fwd_bwd_samples_per_sec", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16" if dt_name == "bf16" else "f16x3", "data": "synthetic",
            "config": {"workload": f"tri-modal contrastive loss fwd+bwd, global batch {b}, dim {d}, "
                                   + ("bf16 embeddings in / bf16 gradients out, fp16 tensor-core operands with fp32 "
                                      "accumulation" if dt_name == "bf16" else
                                      "fp32 embeddings and gradients, split fp16 operands (3 tensor-core products per "
                                      "contraction, fp32 parity mode)"),
                       "name": args.workload, "rows_global": b, "rows_per_gpu": main_res["rows_local"], "dim": d,
                       "parallelism": f"row-strip dp{world}",
                       "l2": "no explicit flush: each step streams the 3 G' strips (2*B*B/world bytes each) through "
                             "HBM, far larger than the 126 MB L2" if b >= 8192 else
                             "no explicit flush: this BASELINE config is smaller than L2 by definition"},
            "tflops_algorithmic_per_gpu": tflops_per_gpu,
            "frac_of_bf16_peak_per_gpu": tflops_per_gpu / peaks["bf16_tflops"],
            "frac_of_bf16_sustained_peak_per_gpu": (tflops_per_gpu / peaks["bf16_tflops_sustained"])
            if peaks.get("bf16_tflops_sustained") else None,
            "clocks": main_res["clocks"],
            "e2e": {"value": b / (main_res["ms_e2e"] * 1e-3), "unit": "samples/s", "ms_per_step": main_res["ms_e2e"],
                    "h2d_bytes_per_step": main_res["h2d"], "d2h_bytes_per_step": 12},
            "gpu_launches": main_res["launches"],  # kernels of libsclip.so launched in the timed region, all ranks
            "roofline": {"bound": "tensor", "kernel": "gemm_wide_kernel (backward_gemms)", "achieved": gemm_tf,
                         "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": (gemm_tf / peaks["bf16_tflops"]) if gemm_tf else None,
                         "traffic": tstep,
                         "traffic_note": None if tstep is not None else
                         "the committed ncu --set full capture (profiles/traffic.json) is of the 1-GPU north-star launch; "
                         "no capture exists for this workload / world size (ncu is never run on multi-rank commands)",
                         "algorithmic_flops_per_launch": 12.0 * b * b * d / world,
                         "peak_source": peaks["source"] + ", burst figure",
                         "stage_ms": st, "stage_rates": stage_rates},
            "cpu_baseline": cpu,
            "gpu_eager_baseline": eager,
            "loss": main_res["loss"],
        }
        if also:
            line["also"] = also
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_sweep(args, measure, world, rank, peaks, dist):
    """BASELINE config 5: global batch 4k-128k x dim 512/768/1024 at this world size.  One JSON line per point into
    gpurun_out/sweep_w<N>.jsonl (copied to profiles/ by hand), one summary line on stdout."""
    import torch

    free, _ = torch.cuda.mem_get_info()
    points = []
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"sweep_w{world}.jsonl")
    f = open(path, "w") if rank == 0 else None
    for b in SWEEP_B:  # small batches first; every point starts from a GPU that has idled for a moment (see main)
        for d in SWEEP_D:
            if b % (world * 256) != 0:
                continue
            time.sleep(0.5)
            need = 3 * 2 * (b // world) * b + 40 * b * d  # the three fp16 G' strips + operands / gradients
            if need > 0.8 * free:
                continue
            steps = 3 if b >= 65536 else 5
            r = measure(b, d, "bf16", True, steps, 3, with_e2e=False, with_stages=False)
            tf = 18.0 * b * b * d / world / (r["ms"] * 1e-3) / 1e12
            pt = {"rows_global": b, "dim": d, "n_gpus": world, "ms_per_step": r["ms"], "samples_per_s": b / (r["ms"] * 1e-3),
                  "tflops_algorithmic_per_gpu": tf, "frac_of_bf16_peak_per_gpu": tf / peaks["bf16_tflops"],
                  "sm_mhz": r["clocks"].get("sm_mhz"), "loss": r["loss"][0]}
            points.append(pt)
            if f is not None:
                f.write(json.dumps(pt) + "\n")
                f.flush()
    if f is not None:
        f.close()
        best = max(points, key=lambda p: p["frac_of_bf16_peak_per_gpu"]) if points else None
        print(json.dumps({
            "metric": "contrastive_loss_fwd_bwd_samples_per_sec", "value": best["samples_per_s"] if best else None,
            "unit": "samples/s", "n_gpus": world, "steps": 5, "warmup": 3,
            "ms_per_step": best["ms_per_step"] if best else None, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "BASELINE config 5: global-batch sweep 4k-128k x dim 512/768/1024 (value = the best point)",
                       "points_file": os.path.relpath(path, ROOT)},
            "sweep": points}), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_step(args, world, rank, dev, dist):
    """BASELINE config 4 / SURVEY 8(f3): the loop body of main_pretraining.py:158-177 (DDP wrap :138, AdamW :139, three
    weighted losses, four `.item()` reads per micro-batch, `loss / accumulation_steps`, `.backward()`, optimizer step
    every 4 micro-batches) on the Base configuration (config.py: ViT-B/16, RoBERTa-base-sized text model, AST-base,
    projection 768) with HF-config-initialised random weights (no network: no checkpoints) and synthetic 224^2 images /
    32-token text / 1024 x 128 spectrograms, per-GPU micro-batch 35 (main_pretraining.py:79).  Arms: the reference's own
    tail statements, the fused tail on the local batch, the fused tail on the global batch (negatives from all ranks);
    each in the reference's regime (fp32, gradient all-reduce on every micro-step) and with bf16 autocast + `no_sync`
    on the non-boundary micro-steps.  Reports ms per optimizer step and the share of the loss tail."""
    import contextlib
    import types

    import torch
    import transformers
    from torch.nn.parallel import DistributedDataParallel as DDP
    from transformers import (ASTConfig, ASTModel, CLIPVisionConfig, CLIPVisionModel, RobertaConfig, RobertaModel)

    from synergy_clip_b200.model import Tri_CLIP

    transformers.CLIPVisionModel.from_pretrained = staticmethod(lambda path: CLIPVisionModel(CLIPVisionConfig(
        hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, image_size=224,
        patch_size=16)))
    transformers.AutoModel.from_pretrained = staticmethod(lambda path: RobertaModel(RobertaConfig(
        hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, vocab_size=50265,
        max_position_embeddings=514)))
    transformers.ASTModel.from_pretrained = staticmethod(lambda path: ASTModel(ASTConfig(
        hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, max_length=1024,
        num_mel_bins=128, patch_size=16)))

    class _Sub:
        output_attentions = False
        output_hidden_states = False
        hidden_size = 768

    cfg = types.SimpleNamespace(vision_config=_Sub, text_config=_Sub, audio_config=_Sub, projection_dim=768,
                                logit_scale_init_value=LOGIT_SCALE_INIT, return_dict=False, is_PT=True,
                                return_logits=False, return_lhs=False)
    micro, accum = 35, 4  # main_pretraining.py:79-80 (Base)
    steps, warmup = min(args.steps, 3), 1
    torch.manual_seed(17)
    model = Tri_CLIP(cfg).to(dev)
    ddp = DDP(model, device_ids=[dev.index]) if world > 1 else model
    opt = torch.optim.AdamW(ddp.parameters(), lr=5e-6)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    batch = dict(pixel_values=torch.randn(micro, 3, 224, 224, device=dev, generator=g),
                 input_ids=torch.randint(3, 50000, (micro, 32), device=dev, generator=g),
                 att_mask=torch.ones(micro, 32, dtype=torch.long, device=dev),
                 input_values=torch.randn(micro, 1024, 128, device=dev, generator=g))

    def optimizer_step(autocast, use_no_sync):
        logged = 0.0
        opt.zero_grad()
        for i in range(accum):
            sync_ctx = ddp.no_sync() if (use_no_sync and world > 1 and i + 1 < accum) else contextlib.nullcontext()
            amp = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
            with sync_ctx:
                with amp:
                    out = ddp(**batch)
                it, ta, ai = out[0] * 1.0, out[1] * 1.0, out[2] * 1.0
                loss = it + ta + ai
                logged += loss.item() + it.item() + ta.item() + ai.item()  # the four reads of :169-170
                (loss / accum).backward()
        opt.step()
        return logged

    def time_arm(env, autocast, use_no_sync):
        for k in ("SCLIP_REFERENCE_TAIL", "SCLIP_GLOBAL_BATCH", "SCLIP_FUSED_PROJECTION"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(warmup):
            optimizer_step(autocast, use_no_sync)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            optimizer_step(autocast, use_no_sync)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    arms = {}
    for regime, (autocast, ns) in (("fp32_allreduce_every_microstep", (False, False)), ("bf16_autocast_no_sync", (True, True))):
        arms[regime] = {
            "reference_tail": time_arm({"SCLIP_REFERENCE_TAIL": "1"}, autocast, ns),
            "fused_tail_local_batch": time_arm({}, autocast, ns),
            "fused_tail_global_batch": time_arm({"SCLIP_GLOBAL_BATCH": "1"}, autocast, ns) if world > 1 else None,
            "fused_projection_and_tail": time_arm({"SCLIP_FUSED_PROJECTION": "1"}, autocast, ns),
        }
    # the loss tail alone at the shapes this step feeds it (local 35 x 768; global 35 * world x 768), fwd + bwd
    from synergy_clip_b200 import ops

    def tail_ms(rows, fused):
        e = [torch.randn(rows, 768, device=dev, generator=g).requires_grad_(True) for _ in range(3)]
        t = [torch.tensor(LOGIT_SCALE_INIT, device=dev, requires_grad=True) for _ in range(3)]

        def run():
            if fused:
                for p in (*e, *t):
                    p.grad = None
                a, b, c = ops.fused_tri_contrastive(*e, *t)
                (a + b + c).backward()
            else:
                reference_tail_eager(*e, t)

        for _ in range(5):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 30

    tails = {"reference_eager_local_ms": tail_ms(micro, False), "fused_local_ms": tail_ms(micro, True),
             "reference_eager_global_rows_ms": tail_ms(micro * world, False),
             "fused_global_rows_ms_single_gpu": tail_ms(micro * world, True)}
    if rank == 0:
        base = arms["fp32_allreduce_every_microstep"]["reference_tail"]
        best = arms["bf16_autocast_no_sync"]["fused_tail_local_batch"]
        print(json.dumps({
            "metric": "pretraining_step_samples_per_sec", "value": micro * accum * world / (best * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": best, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE config 4: Synergy-CLIP Base pre-training step (ViT-B/16 + RoBERTa-base-sized "
                                   "text + AST-base encoders, random init, fused loss tail), synthetic 224^2 images / 32 "
                                   "tokens / 1024x128 spectrograms, micro-batch 35 per GPU x 4 accumulation steps, DDP + "
                                   "AdamW; value = the bf16-autocast + no_sync arm with the fused tail",
                       "micro_batch_per_gpu": micro, "accumulation_steps": accum},
            "ms_per_optimizer_step": arms,
            "speedup_vs_reference_regime": base / best,
            "loss_tail_alone_ms": tails,
            "loss_tail_share_of_microstep": {
                "reference_fp32": tails["reference_eager_local_ms"] / (base / accum),
                "fused_bf16": tails["fused_local_ms"] / (best / accum)},
        }), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
