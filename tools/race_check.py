#!/usr/bin/env python
"""Determinism / ordering check of one ResNet-50 forward + backward through the harness's kernels (1 GPU).

The same weights and the same batch, gradients compared BIT FOR BIT between
  * weight gradients on the main stream and on the side stream (ops.set_wgrad_side_stream),
  * repeated runs of each,
and a digest of every variant printed, so that two processes (TP_PDL=1 / TP_PDL=0: programmatic dependent launch on / off)
can be diffed.  Any difference is an ordering bug: none of these switches changes what is computed.
"""
import hashlib
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch


def main():
    from turboprune_b200 import ops, fused_norm
    from turboprune_b200.utils import config as C
    from turboprune_b200.harness_definitions.standard_pruning_harness import PruningHarness
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cfg = C.compose("synthetic_rn50_erk80", [f"dataset_params.total_batch_size={batch}", "dataset_params.synthetic_steps_per_epoch=2",
                                             f"experiment_params.base_dir={tempfile.gettempdir()}"], os.path.join(ROOT, "conf_b200"))
    import run_experiment
    h = run_experiment.build_harness(cfg) if hasattr(run_experiment, "build_harness") else None
    if h is None:
        from turboprune_b200.utils import custom_models as cm, pruning_utils as pu
        torch.manual_seed(0)
        model = cm.TorchVisionModel(cfg).to("cuda").train()
        torch.manual_seed(1)
        pu.prune_er_erk(model, 0.2)
        h = PruningHarness(cfg=cfg, gpu_id=0, expt_dir=("race", tempfile.gettempdir()), model=model)
        h._setup_optimizer()
    h.model.train()
    x, t = next(iter(h.train_loader))
    x, t = x.cuda(), t.cuda()
    store = h._grad_store()
    names = [n for n, _ in h.model.named_parameters()]
    params = [p for _, p in h.model.named_parameters()]

    def run(side):
        store.zero()
        h._weight_stager().stage()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = h.criterion(h.model(x), t)
        ops.set_wgrad_side_stream(side)
        try:
            loss.backward()
        finally:
            ops.set_wgrad_side_stream(False)
            ops.join_wgrad(h.device)
            fused_norm.drop_partials()
        torch.cuda.synchronize()
        h._drop_staged()
        return loss.detach().clone(), [p.grad.detach().clone() for p in params]

    ref_loss, ref = run(False)
    bad = 0
    for rep in range(4):
        for side in (False, True):
            loss, g = run(side)
            diff = [(n, float((a - b).abs().max())) for n, a, b in zip(names, ref, g) if not torch.equal(a, b)]
            if diff or not torch.equal(loss, ref_loss):
                bad += 1
                print(f"rep {rep} side={side}: loss {float(loss)} vs {float(ref_loss)}; {len(diff)} parameters differ; first: {diff[:6]}", flush=True)
    flat = torch.cat([g.reshape(-1) for g in ref]).cpu().numpy().tobytes()
    print(f"TP_PDL={os.environ.get('TP_PDL', '0')} batch={batch} loss={float(ref_loss):.9f} grad_digest={hashlib.sha1(flat).hexdigest()}")
    per = {n: hashlib.sha1(g.cpu().numpy().tobytes()).hexdigest()[:12] for n, g in zip(names, ref)}
    out = os.path.join(ROOT, "gpurun_out", f"race_digest_pdl{os.environ.get('TP_PDL', '0')}_b{batch}.txt")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        for n in names:
            f.write(f"{n} {per[n]}\n")
    print("RACE CHECK", "PASS" if bad == 0 else f"FAIL ({bad} variants differ)")


def steps_mode():
    """python tools/race_check.py steps <batch> <nsteps>: whole train steps through PruningHarness.train_step (CUDA-graph
    replay), two harnesses from the same initial state on the same batches, weights compared bit for bit after every step:
    variant A = (side-stream wgrad, PDL) as given by the defaults, variants B = one switch flipped."""
    import copy
    from turboprune_b200 import ops, _cabi
    from turboprune_b200.utils import config as C, custom_models as cm, pruning_utils as pu
    from turboprune_b200.harness_definitions.standard_pruning_harness import PruningHarness
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    fresh = len(sys.argv) > 4 and sys.argv[4] == "fresh"      # a new Philox batch every step (what bench.py trains on)
    nosync = "nosync" in sys.argv
    lib = _cabi.load()

    def make(side, graph=True):
        cfg = C.compose("synthetic_rn50_erk80", [f"dataset_params.total_batch_size={batch}", f"dataset_params.synthetic_steps_per_epoch={nsteps}",
                                                 "optimizer_params.weight_decay=1e-4", f"+dataset_params.synthetic_fresh={'true' if fresh else 'false'}",
                                                 f"+experiment_params.wgrad_side_stream={'true' if side else 'false'}",
                                                 f"+experiment_params.cuda_graph={'true' if graph else 'false'}",
                                                 f"experiment_params.base_dir={tempfile.gettempdir()}"], os.path.join(ROOT, "conf_b200"))
        torch.manual_seed(0)
        model = cm.TorchVisionModel(cfg)
        torch.manual_seed(1)
        pu.prune_er_erk(model, 0.2)
        h = PruningHarness(cfg=cfg, gpu_id=0, expt_dir=("race", tempfile.gettempdir()), model=model)
        h._setup_optimizer(); h._setup_scheduler(1)
        h.model.train()
        return h

    def trajectory(side, pdl, graph=True):
        lib.tp_set_pdl(1 if pdl else 0)
        h = make(side, graph)
        snaps, losses = [], []
        for b in h.train_loader:
            losses.append(h.train_step(b)["loss"].clone())
            h.scheduler.step()
            if not nosync:
                torch.cuda.synchronize()
                snaps.append([p.detach().clone() for p in h.model.parameters()])
        if nosync:                         # free-running host (what bench.py does): only the final weights are compared
            torch.cuda.synchronize()
            snaps.append([p.detach().clone() for p in h.model.parameters()])
        names = [n for n, _ in h.model.named_parameters()]
        del h
        torch.cuda.empty_cache()
        lib.tp_set_pdl(0)
        return names, snaps, [float(l) for l in losses]

    names, ref, ref_loss = trajectory(False, False, graph=False)
    print("reference (eager, one stream, no PDL) loss mean:", sum(ref_loss) / len(ref_loss), "last:", ref_loss[-4:], flush=True)
    bad = 0
    for label, side, pdl, graph in (("eager again", False, False, False), ("graph", False, False, True), ("graph+side", True, False, True),
                                    ("graph+pdl", False, True, True), ("graph+side+pdl", True, True, True), ("eager+side+pdl", True, True, False)):
        _, snaps, losses = trajectory(side, pdl, graph)
        first = None
        for s_, (a, b) in enumerate(zip(ref, snaps)):
            diff = [(n, float((x - y).abs().max())) for n, x, y in zip(names, a, b) if not torch.equal(x, y)]
            if diff:
                first = (s_, len(diff), diff[:5])
                break
        if first:
            bad += 1
        print(f"{label:>16}: loss mean {sum(losses) / len(losses):.6f}  ->", "identical weights after every step" if first is None
              else f"FIRST DIFFERENCE after step {first[0]}: {first[1]} parameters, e.g. {first[2]}", flush=True)
    print("RACE CHECK (steps)", "PASS" if bad == 0 else f"FAIL ({bad} variants differ)")


if __name__ == "__main__":
    steps_mode() if len(sys.argv) > 1 and sys.argv[1] == "steps" else main()
