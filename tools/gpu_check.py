#!/usr/bin/env python
"""One-shot GPU bring-up checks: every kernel family against a reference, each group in its own
subprocess (a trapped kernel poisons the CUDA context) and under a timeout.  Writes a report to
gpurun_out/gpu_check.txt.  Usage: python tools/gpu_check.py [group ...]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GROUPS = ["prune", "optim", "gemm", "conv", "dgrad", "wgrad", "layers", "bn", "model"]


def _rel(a, b):
    import torch
    a = a.detach().float(); b = b.detach().float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def g_prune():
    import numpy as np, torch
    from turboprune_b200 import ops, _cabi
    from oracle import prune as P
    dev = "cuda"
    rng = np.random.RandomState(0)
    def run(ws, ms, k, gs=None, kind=0, tag=""):
        tw = [torch.from_numpy(w).to(dev) for w in ws]; tm = [torch.from_numpy(m).to(dev) for m in ms]
        tg = None if gs is None else [torch.from_numpy(g).to(dev) for g in gs]
        torch.cuda.synchronize(); t = time.time()
        outs, thr, info = ops.topk_threshold_mask(tw, tm, k, gs=tg, kind=kind)
        torch.cuda.synchronize(); dt = time.time() - t
        sc = P.layer_scores(ws, ms, gs, kind)
        ref_thr = P.kth_smallest(np.concatenate([s.ravel() for s in sc]), k)
        ref = [P.apply_threshold(s, ref_thr) for s in sc]
        ok_thr = (np.float32(thr.item()).view(np.uint32) == np.float32(ref_thr).view(np.uint32)) or (np.isnan(ref_thr) and np.isnan(thr.item()))
        ok_m = all(np.array_equal(o.cpu().numpy(), r) for o, r in zip(outs, ref))
        print(f"  prune[{tag}] k={k} thr={thr.item():.6g} ref={ref_thr:.6g} thr_ok={bool(ok_thr)} mask_ok={ok_m} info={info} t={dt*1e3:.2f}ms")
        return ok_thr and ok_m
    ok = True
    sizes = [1000, 4096 * 3 + 17, 300000, 2_000_003]
    ws = [rng.randn(n).astype(np.float32) * 0.05 for n in sizes]
    ms = [np.ones(n, np.float32) for n in sizes]
    N = sum(sizes)
    ok &= run(ws, ms, int(0.2 * N), tag="mag")
    ok &= run(ws, ms, 1, tag="k=1")
    ok &= run(ws, ms, N, tag="k=N")
    ms2 = [(rng.rand(n) < 0.5).astype(np.float32) for n in sizes]
    ok &= run(ws, ms2, int(0.6 * N), tag="ties-zero")
    ok &= run(ws, ms2, int(0.3 * N), tag="thr-in-zeros")
    gs = [rng.randn(n).astype(np.float32) * 1e-3 for n in sizes]
    ok &= run(ws, ms2, int(0.7 * N), gs=gs, kind=1, tag="snip")
    ok &= run(ws, ms2, int(0.7 * N), gs=gs, kind=2, tag="synflow")
    we = [np.full(n, 0.25, np.float32) for n in sizes]
    ok &= run(we, ms, int(0.5 * N), tag="all-equal(fallback)")
    wn = [w.copy() for w in ws]; wn[2][:5000] = np.nan
    ok &= run(wn, ms, N - 100, tag="nan-thr")
    ok &= run(wn, ms, int(0.5 * N), tag="nan-present")
    try:
        ops.topk_threshold_mask([torch.from_numpy(ws[0]).to(dev)], [torch.from_numpy(ms[0]).to(dev)], 0)
        print("  k=0 did NOT raise"); ok = False
    except RuntimeError as e:
        print("  k=0 raises:", str(e)[:60])
    # count zeros
    cz = ops.count_zeros([torch.from_numpy(m).to(dev) for m in ms2]).tolist()
    ok &= cz[:-1] == [int((m == 0).sum()) for m in ms2] and cz[-1] == sum(cz[:-1])
    print("  count_zeros ok:", cz[-1])
    # big: RN50-sized
    n = 25_502_912
    w = torch.randn(n, device=dev) * 0.03; m = torch.ones(n, device=dev)
    k = int(0.2 * n)
    for _ in range(2):
        outs, thr, info = ops.topk_threshold_mask([w], [m], k)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); outs, thr, info = ops.topk_threshold_mask([w], [m], k); e1.record(); torch.cuda.synchronize()
    ms_ = e0.elapsed_time(e1)
    ref = torch.kthvalue(w.abs(), k)[0]
    print(f"  RN50-size: {ms_:.3f} ms -> {12*n/ms_/1e6:.1f} GB/s info={info} thr_ok={bool(ref == thr)} zeros={int((outs[0]==0).sum())} (k={k})")
    ok &= bool(ref == thr) and int((outs[0] == 0).sum()) == int((w.abs() <= ref).sum())
    return ok


def g_optim():
    import numpy as np, torch
    from turboprune_b200 import ops
    from oracle.train import sgd_momentum_step
    dev = "cuda"; rng = np.random.RandomState(1)
    sizes = [5, 4096, 100001]
    ws = [rng.randn(n).astype(np.float32) for n in sizes]; gs = [rng.randn(n).astype(np.float32) for n in sizes]
    tw = [torch.from_numpy(w.copy()).to(dev) for w in ws]; tg = [torch.from_numpy(g).to(dev) for g in gs]
    tb = [torch.zeros_like(w) for w in tw]
    lr = torch.tensor(0.1, device=dev)
    bufs = [None] * 3
    ok = True
    for step in range(3):
        ops.sgd_momentum_step(tw, tg, tb, lr, 0.9, 5e-4, step == 0)
        for i in range(3):
            ws[i], bufs[i] = sgd_momentum_step(ws[i], gs[i], bufs[i], 0.1, 0.9, 5e-4, step == 0)
        err = max(float(np.abs(tw[i].cpu().numpy() - ws[i]).max()) for i in range(3))
        print(f"  sgd step {step} max abs err {err:.3e}")
        ok &= err < 1e-6
    return ok


def _conv_case(n, h, w, cin, cout, r, s, stride, pad, bias=False, check=("f", "d", "w")):
    import torch, torch.nn.functional as F
    from turboprune_b200 import ops
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(n * 7 + cin + cout + r)
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    wt = (torch.randn(cout, cin, r, s, generator=g) / (cin * r * s) ** 0.5).to(dev)
    mk = (torch.rand(cout, cin, r, s, generator=g) < 0.4).float().to(dev)
    b = torch.randn(cout, generator=g).to(dev) if bias else None
    xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_("d" in check)
    wp = wt.clone().requires_grad_(True)
    bp = b.clone().requires_grad_(True) if bias else None
    y = ops.masked_conv2d(xb, wp, mk, bp, (stride, stride), (pad, pad))
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    xr = xb.detach().float().requires_grad_(True)
    wr = (mk * wt).to(torch.bfloat16).float().requires_grad_(True)
    yr = F.conv2d(xr, wr, b, stride, pad)
    res = {}
    res["f"] = _rel(y, yr)
    print(f"  [fwd] n{n} {h}x{w} c{cin}->{cout} k{r} s{stride} p{pad}: rel={res['f']:.3e}", flush=True)
    dy = torch.randn(y.shape, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y.backward(dy)
    yr.backward(dy.float())
    if "d" in check:
        res["d"] = _rel(xb.grad, xr.grad)
    res["w"] = _rel(wp.grad, mk * wr.grad)
    masked_zero = bool((wp.grad[mk == 0] == 0).all())
    if bias:
        res["b"] = _rel(bp.grad, dy.float().sum((0, 2, 3)))
    ok = all(v < 2e-2 for v in res.values()) and masked_zero
    print(f"  conv n{n} {h}x{w} c{cin}->{cout} k{r} s{stride} p{pad} bias={bias}: " +
          " ".join(f"{k}={v:.2e}" for k, v in res.items()) + f" masked_grad_zero={masked_zero} {'OK' if ok else 'FAIL'}")
    return ok


def g_gemm():
    ok = True
    ok &= _conv_case(2, 8, 8, 64, 64, 1, 1, 1, 0)          # 128 pixels: exactly one tile
    ok &= _conv_case(3, 7, 7, 128, 256, 1, 1, 1, 0)        # partial M tile, 2 k-blocks
    ok &= _conv_case(4, 14, 14, 256, 1024, 1, 1, 1, 0)     # several tiles
    return ok


def g_conv():
    ok = True
    ok &= _conv_case(2, 8, 8, 64, 64, 3, 3, 1, 1, check=("f",))
    ok &= _conv_case(2, 14, 14, 128, 128, 3, 3, 1, 1, check=("f",))
    ok &= _conv_case(2, 14, 14, 128, 128, 3, 3, 2, 1, check=("f",))
    ok &= _conv_case(2, 14, 14, 256, 512, 1, 1, 2, 0, check=("f",))
    return ok


def g_dgrad():
    ok = True
    ok &= _conv_case(2, 8, 8, 64, 64, 3, 3, 1, 1)
    ok &= _conv_case(2, 14, 14, 128, 128, 3, 3, 2, 1)
    ok &= _conv_case(2, 14, 14, 256, 512, 1, 1, 2, 0)
    ok &= _conv_case(3, 7, 7, 512, 512, 3, 3, 1, 1)
    return ok


def g_wgrad():
    ok = True
    ok &= _conv_case(8, 28, 28, 128, 128, 3, 3, 1, 1, bias=True)
    ok &= _conv_case(4, 56, 56, 64, 256, 1, 1, 1, 0)
    ok &= _conv_case(2, 32, 32, 3, 64, 3, 3, 1, 1, bias=True, check=("f",))    # stem (CIFAR / VGG)
    ok &= _conv_case(2, 64, 64, 3, 64, 7, 7, 2, 3, check=("f",))               # stem (ImageNet)
    return ok


def g_layers():
    import torch
    from turboprune_b200.utils.mask_layers import ConvMask, Conv1dMask, LinearMask
    import oracle.mask_ops as R
    torch.manual_seed(0)
    dev = "cuda"; ok = True
    fc = Conv1dMask(2048, 1000, bias=True).to(dev)
    fc.set_er_mask(0.3)
    x = torch.randn(64, 2048, device=dev, dtype=torch.bfloat16, requires_grad=True)
    y = fc(x); dy = torch.randn_like(y); y.backward(dy)
    yr = R.masked_conv1d_k1(x.detach().cpu(), fc.weight.detach().cpu(), fc.mask.cpu(), fc.bias.detach().cpu(), bf16_operands=True)
    gx, gw, gb = R.masked_linear_grads(x.detach().cpu(), fc.weight.detach().cpu()[:, :, 0], fc.mask.cpu()[:, :, 0], dy.cpu(), True, True)
    e = dict(f=_rel(y.cpu(), yr), d=_rel(x.grad.cpu(), gx), w=_rel(fc.weight.grad.cpu()[:, :, 0], gw), b=_rel(fc.bias.grad.cpu(), gb))
    print("  Conv1dMask 2048->1000:", e); ok &= all(v < 2e-2 for v in e.values())
    fc = Conv1dMask(512, 10, bias=True).to(dev)
    x = torch.randn(96, 512, device=dev, dtype=torch.bfloat16, requires_grad=True)
    y = fc(x); dy = torch.randn_like(y); y.backward(dy)
    yr = R.masked_conv1d_k1(x.detach().cpu(), fc.weight.detach().cpu(), fc.mask.cpu(), fc.bias.detach().cpu(), bf16_operands=True)
    gx, gw, gb = R.masked_linear_grads(x.detach().cpu(), fc.weight.detach().cpu()[:, :, 0], fc.mask.cpu()[:, :, 0], dy.cpu(), True, True)
    e = dict(f=_rel(y.cpu(), yr), d=_rel(x.grad.cpu(), gx), w=_rel(fc.weight.grad.cpu()[:, :, 0], gw), b=_rel(fc.bias.grad.cpu(), gb))
    print("  Conv1dMask 512->10:", e); ok &= all(v < 2e-2 for v in e.values())
    lin = LinearMask(in_features=384, out_features=1152, bias=True).to(dev)
    x = torch.randn(4, 197, 384, device=dev, dtype=torch.bfloat16, requires_grad=True)
    y = lin(x); dy = torch.randn_like(y); y.backward(dy)
    yr = R.masked_linear(x.detach().cpu(), lin.weight.detach().cpu(), lin.mask.cpu(), lin.bias.detach().cpu(), bf16_operands=True)
    gx, gw, gb = R.masked_linear_grads(x.detach().cpu(), lin.weight.detach().cpu(), lin.mask.cpu(), dy.cpu(), True, True)
    e = dict(f=_rel(y.cpu(), yr), d=_rel(x.grad.cpu(), gx), w=_rel(lin.weight.grad.cpu(), gw), b=_rel(lin.bias.grad.cpu(), gb))
    print("  LinearMask 384->1152 (197 tokens):", e); ok &= all(v < 2e-2 for v in e.values())
    return ok


def g_bn():
    import torch, torch.nn.functional as F
    from turboprune_b200.fused_norm import BatchNorm2dB200
    dev = "cuda"; ok = True
    for (n, c, h, w, relu, res) in [(4, 64, 9, 7, True, False), (8, 256, 14, 14, True, True), (3, 2048, 7, 7, False, False),
                                    (64, 64, 56, 56, True, False), (2, 192, 5, 5, False, True)]:
        g = torch.Generator().manual_seed(c + n)
        x = (torch.randn(n, c, h, w, generator=g) * 1.7 + 0.3).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        r = torch.randn(n, c, h, w, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if res else None
        bn = BatchNorm2dB200(c).to(dev); ref = torch.nn.BatchNorm2d(c).to(dev)
        with torch.no_grad():
            bn.weight.copy_(torch.rand(c, generator=g) + 0.5); bn.bias.copy_(torch.randn(c, generator=g) * 0.1)
        ref.load_state_dict(bn.state_dict())
        xa = x.clone().requires_grad_(True); xb = x.clone().float().requires_grad_(True)
        ra = r.clone().requires_grad_(True) if res else None; rb = r.clone().float().requires_grad_(True) if res else None
        z = bn(xa, residual=ra, relu=relu)
        zr = ref(xb)
        if res: zr = zr + rb
        if relu: zr = torch.relu(zr)
        dz = torch.randn(z.shape, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        z.backward(dz); zr.backward(dz.float())
        e = dict(f=_rel(z, zr), dx=_rel(xa.grad, xb.grad), dw=_rel(bn.weight.grad, ref.weight.grad), db=_rel(bn.bias.grad, ref.bias.grad),
                 rm=_rel(bn.running_mean, ref.running_mean), rv=_rel(bn.running_var, ref.running_var))
        if res: e["dres"] = _rel(ra.grad, rb.grad)
        good = all(v < 1e-2 for v in e.values()) and int(bn.num_batches_tracked) == 1
        ok &= good
        print(f"  bn n{n} c{c} {h}x{w} relu={relu} res={res}: " + " ".join(f"{k}={v:.1e}" for k, v in e.items()) + (" OK" if good else " FAIL"))
        if c % 8 == 0 and h >= 5:
            from turboprune_b200.fused_norm import MaxPool2dB200
            for (k_, s_, p_) in ((3, 2, 1), (2, 2, 0)):
                xa2 = x.clone().requires_grad_(True); xb2 = x.clone().float().requires_grad_(True)
                ya = MaxPool2dB200(k_, s_, p_)(xa2); yb = torch.nn.functional.max_pool2d(xb2, k_, s_, p_)
                dyy = torch.randn(ya.shape, generator=g).to(dev).to(torch.bfloat16)
                ya.backward(dyy); yb.backward(dyy.float())
                mp_ok = torch.equal(ya.float(), yb) and _rel(xa2.grad, xb2.grad) < 1e-2
                ok &= mp_ok
                print(f"    maxpool k{k_} s{s_} p{p_}: fwd exact={torch.equal(ya.float(), yb)} bwd rel={_rel(xa2.grad, xb2.grad):.1e} {'OK' if mp_ok else 'FAIL'}")
        bn.eval(); ref.eval()
        with torch.no_grad():
            ze = bn(x, relu=relu); zre = ref(x.float()); zre = torch.relu(zre) if relu else zre
        ok &= _rel(ze, zre) < 1e-2
    return ok


def g_model():
    import copy, torch
    import oracle.model as om
    from oracle.train import train_step
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from refshim import make_cfg
    from turboprune_b200.utils import custom_models as cm, pruning_utils as pu
    ok = True
    for name, ds, shape, ncls in [("resnet18", "cifar10", (32, 3, 32, 32), 10), ("resnet50", "imagenet", (8, 3, 224, 224), 1000)]:
        torch.manual_seed(0)
        ref = om.build(name, ds)
        torch.manual_seed(0)
        mine = cm.TorchVisionModel(make_cfg(name, ds))
        same = all(torch.equal(a, b) for a, b in zip(ref.state_dict().values(), mine.model.state_dict().values()))
        torch.manual_seed(1)
        pu.prune_er_erk(mine, 0.2)
        ref.load_state_dict(mine.model.state_dict())
        mine = mine.cuda()
        print(f"  {name}: init identical={same} sparsity={mine.get_overall_sparsity():.2f}%")
        g = torch.Generator().manual_seed(5)
        x = torch.randn(*shape, generator=g); t = torch.randint(0, ncls, (shape[0],), generator=g)
        o_ref = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
        o_mine = torch.optim.SGD(mine.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
        ref.train(); mine.train()
        for step in range(2):
            l_ref, _ = train_step(ref, o_ref, x, t)
            l_mine, _ = train_step(mine, o_mine, x.cuda(), t.cuda(), device_type="cuda")
            rel = abs(l_ref - l_mine) / abs(l_ref)
            print(f"    step {step}: loss ref={l_ref:.6f} ours={l_mine:.6f} rel={rel:.2e}")
            ok &= rel < (1e-3 if step == 0 else 5e-2)
        gerr = []
        for (n1, p1), (n2, p2) in zip(ref.named_parameters(), mine.model.named_parameters()):
            gerr.append((_rel(p2.grad.cpu(), p1.grad), n1))
        gerr.sort(reverse=True)
        print("    worst grad rel errs:", [(f"{e:.2e}", n) for e, n in gerr[:4]])
    return ok


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        import torch
        torch.manual_seed(0)
        fn = globals()["g_" + sys.argv[2]]
        try:
            ok = fn()
        except Exception:
            import traceback; traceback.print_exc(); ok = False
        print(f"GROUP {sys.argv[2]}: {'PASS' if ok else 'FAIL'}")
        sys.exit(0 if ok else 1)
    groups = sys.argv[1:] or GROUPS
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    report = []
    for gname in groups:
        t = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", gname], cwd=ROOT,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=420)
            out, code = r.stdout, r.returncode
        except subprocess.TimeoutExpired as e:
            out, code = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), -9
            out += "\nTIMEOUT"
        report.append(f"===== {gname} (exit {code}, {time.time()-t:.1f}s) =====\n{out}")
        print(report[-1], flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.txt"), "w") as f:
        f.write("\n".join(report))


if __name__ == "__main__":
    main()
