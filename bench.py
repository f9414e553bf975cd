#!/usr/bin/env python
"""Headline benchmark: images/sec of the masked data-parallel train step, ResNet-50 @ 80 % unstructured
(ERK) sparsity, ImageNet-shaped synthetic data, bf16 autocast — BASELINE.json's metric and config.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (the reference's CPU path = oracle port, host cores)

One "step" = one pass of the hot path over one batch: H2D (e2e only) -> zero_grad -> autocast forward through
the sm_100a masked-conv kernels -> CE loss -> backward (dgrad/wgrad kernels, mask fused in wgrad) -> P2P gradient
mean over NVLink (N>1) -> fused SGD.  Nothing is skipped in the timed region.

Output: ONE JSON line (see the task contract): value = whole-job images/s with inputs resident in HBM,
e2e = the same through the public API with pinned-host inputs copied every step and the loss read back,
roofline = the masked implicit-GEMM kernels' achieved TFLOP/s (CUDA events on the launching stream, live in the
timed region) against the measured sustained bf16 peak, cpu_baseline = the oracle port timed on host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "images_per_sec_resnet50_erk80_train_step"
GFLOP_PER_IMG = 24.30           # fwd 8.178 + dgrad 7.942 + wgrad 8.178 (masked layers, dense)
ROOFLINE_IMG_S = 39.8e3         # per-layer max(tensor, HBM) masked-GEMM roofline per GPU


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.nv = []            # in-process NVML samples (sm MHz, max MHz, reasons bitmask): every 10 ms, so even a
        self._stop = False      # 0.2 s timed region (8 GPUs x batch 64) is sampled; nvidia-smi -lms stays as the fallback

    def _nvml_loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop:
                self.nv.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), mx,
                                int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
                time.sleep(0.01)
        except Exception:
            pass

    def start(self):
        try:
            self.tn = threading.Thread(target=self._nvml_loop, daemon=True); self.tn.start()
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.nv:
            # NVML throttle-reason bits: 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal_slowdown, 0x40 hw_thermal_slowdown
            bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            sm_sorted = sorted(v[0] for v in self.nv)
            loaded = sm_sorted[len(sm_sorted) // 3:] or sm_sorted
            allbits = 0
            for v in self.nv:
                allbits |= v[2]
            return {"sm_mhz": float(statistics.median(loaded)), "sm_max_mhz": float(self.nv[0][1]),
                    "reasons": sorted(n for b, n in bits.items() if allbits & b), "samples": len(self.nv), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm_sorted = sorted(sm)
        loaded = sm_sorted[len(sm_sorted) // 3:] or sm_sorted       # drop idle samples at the edges
        return {"sm_mhz": statistics.median(loaded) if loaded else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def triangular_lr(total_steps, warmup_fraction=0.2):
    """LR multiplier schedule of the reference (utils/schedulers.py:79-117): interp [0.2, 1, 0]."""
    import numpy as np
    return np.interp(np.arange(1 + total_steps), [0, int(warmup_fraction * total_steps), total_steps], [0.2, 1, 0])


def cpu_train_step_rate(batch, steps, warmup, threads=None):
    """The reference's CPU path (oracle port): RN50 ERK-80 train step, bf16 autocast, on host cores."""
    import torch
    import oracle.model as om
    from oracle import prune as OP
    from oracle.train import train_step
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = om.build("resnet50", "imagenet")
    torch.manual_seed(1)
    probs = OP.erk_keep_probabilities([tuple(m.weight.shape) for _, m in om.masked_layers(net)], 0.2)
    om.set_er_masks(net, probs)
    opt = torch.optim.SGD(net.parameters(), lr=0.2, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(batch, 3, 224, 224, generator=g); t = torch.randint(0, 1000, (batch,), generator=g)
    net.train()
    for _ in range(warmup):
        train_step(net, opt, x, t)
    t0 = time.perf_counter()
    for _ in range(steps):
        train_step(net, opt, x, t)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 16
    steps = max(1, min(args.steps, 40))
    rate, s_per_step, threads = cpu_train_step_rate(batch, steps, max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "resnet50 imagenet-shape ERK-80% masked train step (reference CPU path, oracle port)",
                   "global_batch": batch, "sample": f"batch {batch} per step on host cores"},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} steps of batch {batch} (torch CPU, bf16 autocast), os.cpu_count()={os.cpu_count()}"},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--global-batch", type=int, default=512, help="reference semantics: total_batch_size split over ranks")
    ap.add_argument("--per-gpu-batch", type=int, default=0, help="override: fixed per-GPU batch (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-topk", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the train step into a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from refshim import make_cfg
    from turboprune_b200 import ops
    from turboprune_b200.optim import FusedSGD
    from turboprune_b200.utils import custom_models as cm, pruning_utils as pu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(1, args.warmup)
    K = max(1, args.steps)
    B = args.per_gpu_batch or max(1, args.global_batch // world)
    scaling = "weak" if args.per_gpu_batch else "strong"

    # ---- model: seed-0 ResNet-50, ERK-80 % Bernoulli masks (identical on every rank by construction) ----
    torch.manual_seed(0)
    model = cm.TorchVisionModel(make_cfg("resnet50", "imagenet", precision="bfloat16"))
    torch.manual_seed(1)
    pu.prune_er_erk(model, 0.2)
    model = model.to(dev).train()
    sparsity = model.get_overall_sparsity()
    use_graph = not args.no_graph
    opt = FusedSGD(model.parameters(), lr=0.2, momentum=0.9, weight_decay=1e-4, capturable=use_graph)
    sched_tab = triangular_lr(10 * (W + K) * 3)
    reducer = None
    if world > 1:
        from turboprune_b200.grad_exchange import P2PGradReducer
        reducer = P2PGradReducer(list(model.parameters()))

    gen = torch.Generator(device=dev).manual_seed(1000 * 0 + rank)
    pool = [(torch.randn(B, 3, 224, 224, device=dev, generator=gen).contiguous(memory_format=torch.channels_last),
             torch.randint(0, 1000, (B,), device=dev, generator=gen)) for _ in range(2)]
    step_idx = [0]
    loss_acc = torch.zeros((), device=dev)

    def set_lr():
        for g in opt.param_groups:
            g["lr"] = 0.2 * float(sched_tab[min(step_idx[0], len(sched_tab) - 1)])
        step_idx[0] += 1
        if use_graph:
            opt.sync_lr()                      # device scalar: the captured step reads it, nothing is baked in

    from turboprune_b200.grad_exchange import GradArena
    arena = reducer if reducer is not None else GradArena(list(model.parameters()))
    from turboprune_b200 import ops as _ops
    from turboprune_b200.utils.mask_layers import MASKED_LAYER_TYPES
    stager = _ops.WeightStager([m for m in model.modules() if isinstance(m, MASKED_LAYER_TYPES)])

    def step_body(x, t):
        arena.zero()                           # one memset; grads live in persistent slots (stable pointers)
        stager.stage()                         # bf16(mask*w) operands of all 54 layers: one launch
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(x)
            loss = torch.nn.functional.cross_entropy(out, t)
        loss.backward()
        if reducer is not None:
            reducer.reduce()
        opt.step()
        return loss

    def eager_step(x, t):
        set_lr()
        return step_body(x, t)

    step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(nsteps, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nsteps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up (eager): allocates workspaces, sets kernel attributes, fills momentum buffers ----
    for i in range(W):
        step(*pool[i % 2])
    launches_per_step = 0
    if True:
        l0 = ops.launch_count(); step(*pool[0]); launches_per_step = ops.launch_count() - l0

    # ---- capture ONE whole train step (fwd + bwd + P2P reduce + SGD) into a CUDA graph ----
    graph = None
    if use_graph:
        static_x, static_t = torch.empty_like(pool[0][0]), torch.empty_like(pool[0][1])
        static_x.copy_(pool[0][0]); static_t.copy_(pool[0][1])
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                eager_step(static_x, static_t)
        torch.cuda.current_stream(dev).wait_stream(side)
        barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            static_loss = step_body(static_x, static_t)

        def graph_step(x, t):
            static_x.copy_(x, non_blocking=True); static_t.copy_(t, non_blocking=True)
            set_lr()
            graph.replay()
            return static_loss
        step = graph_step
        for i in range(2):
            step(*pool[i % 2])

    # ---- device-resident run (value) ----
    clocks = ClockSampler(local_rank); clocks.start()

    def dev_step(i):
        loss_acc.add_(step(*pool[i % 2]).detach())
    ms_total = timed(K, dev_step)
    launches = launches_per_step * K
    clk = clocks.stop()
    if reducer is not None:
        reducer.check_status()
    img_s = world * B * K / (ms_total / 1e3)

    # ---- per-kernel timing of the masked GEMMs: CUDA events around every C-ABI conv call on the launching stream
    # (eager steps of the same workload — events cannot be read back from inside a replayed graph) ----
    timer = ops.KernelTimer()
    KT = min(K, 5)
    ops.set_timer(timer)
    ms_eager = timed(KT, lambda i: eager_step(*pool[i % 2]))
    ops.set_timer(None)
    tot = timer.totals()
    gemm_ms = sum(v[0] for v in tot.values()) * (K / KT)
    pk = peaks()
    flops = GFLOP_PER_IMG * 1e9 * B * K
    achieved_tf = flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    # Algorithmic bytes of the masked convs/linears (bf16 activations in + out per op):
    # fprop x+y, dgrad dy+dx (the stem has no dgrad), wgrad x+dy.  Shapes taken from the live model.
    io = {}
    hooks = []
    for name, m in model._masked():
        hooks.append(m.register_forward_hook(lambda mod, inp, out, name=name: io.__setitem__(name, (inp[0].numel(), (out[0] if isinstance(out, tuple) else out).numel()))))
    model.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        model(pool[0][0][:2])
    model.train()
    for h in hooks:
        h.remove()
    first = next(iter(io))
    per_img = sum(3 * (a + b) for a, b in io.values()) - sum(io[first])          # elements per 2 images
    alg_bytes_step = 2.0 * per_img / 2 * B
    achieved_gbs = alg_bytes_step * K / (gemm_ms / 1e3) / 1e9 if gemm_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_igemm_traffic.json")
    if os.path.isfile(tpath) and B == 512:
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]
    n_launch = sum(v[2] for v in tot.values()) / KT
    roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": pk["hbm"], "unit": "GB/s",
                "frac": achieved_gbs / pk["hbm"], "traffic": traffic,
                "traffic_note": "dram read+write per igemm kernel launch, ncu over one eager step at B=512 (profiles/r01_igemm_traffic.json)" if traffic else None,
                "peak_source": pk["src"] + " HBM copy bandwidth",
                "kernel": "k_igemm_fwd (fprop+dgrad) / k_igemm_wgrad — masked implicit GEMM, tcgen05 + TMA",
                "why_hbm": "sum over the 54 layers: conv I/O bytes / HBM peak (11.4 ms at B=512) exceeds FLOPs / tensor peak (8.5 ms); 30 of 54 layers are HBM-bound",
                "algorithmic_bytes_per_step": alg_bytes_step, "algorithmic_bytes_per_launch": alg_bytes_step / n_launch,
                "launches_per_step": n_launch,
                "timing": f"CUDA events around each masked-GEMM C-ABI call over {KT} eager steps ({ms_eager / KT:.2f} ms/step eager); step rate from CUDA-graph replay" if use_graph else "CUDA events, eager",
                "ms_per_step_in_kernel": gemm_ms / K,
                "by_op_ms_per_step": {k: v[0] / KT for k, v in tot.items()},
                "share_of_step": gemm_ms / ms_total,
                "tensor": {"achieved_tflops": achieved_tf, "peak_tflops": pk["tf"], "frac": achieved_tf / pk["tf"],
                           "flops_per_step": GFLOP_PER_IMG * 1e9 * B, "peak_source": pk["src"] + " sustained bf16"},
                "frac_of_masked_gemm_roofline_img_s": (img_s / world) / ROOFLINE_IMG_S}

    # ---- end-to-end run: pinned host inputs copied every step, loss read back every step ----
    e2e = None
    if not args.no_e2e:
        hpool = [(torch.randn(B, 3, 224, 224).pin_memory(), torch.randint(0, 1000, (B,)).pin_memory()) for _ in range(2)]
        h2d = hpool[0][0].numel() * 4 + hpool[0][1].numel() * 8

        # double-buffered prefetch: batch i+1 crosses PCIe on a copy stream while step i computes
        copy_stream = torch.cuda.Stream(dev)
        stage = [(torch.empty_like(pool[0][0]), torch.empty_like(pool[0][1])) for _ in range(2)]
        h2d_done = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            j = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[j])
                stage[j][0].copy_(hpool[j][0], non_blocking=True); stage[j][1].copy_(hpool[j][1], non_blocking=True)
                h2d_done[j].record(copy_stream)

        for ev in consumed:
            ev.record(torch.cuda.current_stream(dev))

        def e2e_step(i):
            j = i % 2
            if i == 0:
                prefetch(0)
            torch.cuda.current_stream(dev).wait_event(h2d_done[j])
            prefetch(i + 1)                                   # overlaps with this step's compute
            loss = step(stage[j][0], stage[j][1])
            consumed[j].record(torch.cuda.current_stream(dev))
            return float(loss.item())                         # the reference returns loss.item() every step (base_harness.py:134)
        for i in range(2):
            e2e_step(i)
        torch.cuda.synchronize(dev)
        ms_e2e = timed(K, e2e_step)
        e2e = {"value": world * B * K / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e / K}

    # ---- mask top-k (second half of the metric): prune_mag over the model's 25.5 M masked weights ----
    topk = None
    if not args.no_topk and rank == 0:
        layers = [m for _, m in model._masked()]
        ws = [m.weight.detach() for m in layers]; ms_ = [torch.ones_like(m.mask) for m in layers]
        n = sum(w.numel() for w in ws); k = int((1 - 0.2) * n)
        plan = ops.TopKPlan(ws, ms_)                       # pointer tables marshalled once: the timed call is the C-ABI call
        for _ in range(3):
            plan.run(k)
        reps = 10
        flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
        tms = []
        for _ in range(reps):
            flush.zero_()                                  # 256 MiB write: evicts the 126 MB L2
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); _, _, info = plan.run(k); b.record(); torch.cuda.synchronize(dev)
            tms.append(a.elapsed_time(b))
        tmed = statistics.median(tms)
        gbs = 12.0 * n / (tmed / 1e3) / 1e9
        topk = {"metric": "mask_topk_GBps", "elements": n, "k": k, "algorithmic_bytes": 12 * n, "ms": tmed, "GBps": gbs,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": None},
                "path": info["path"], "candidates": info["candidates"], "l2": "flushed between reps"}
        del flush

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, s_per_step, threads = cpu_train_step_rate(16, 6, 1)
        cpu = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"6 steps of batch 16 of the same workload (oracle port, torch CPU bf16 autocast, {s_per_step:.2f} s/step), "
                         f"os.cpu_count()={os.cpu_count()}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": img_s, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "resnet50 imagenet-shape [B,3,224,224] ERK-80% unstructured masks, SGD(0.9, wd 1e-4), CE loss",
                       "global_batch": B * world, "per_gpu_batch": B, "parallelism": f"dp{world}",
                       "sparsity_percent": sparsity, "cuda_graph": bool(use_graph), "l2": "inputs (308 MB/batch at B=512) and activations exceed the 126 MB L2",
                       "grad_exchange": "none (1 GPU)" if world == 1 else "tp_p2p_allreduce_mask over symmetric memory (NVLink), no NCCL on the data path"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
            "topk": topk, "loss_mean": float(loss_acc.item()) / K,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
