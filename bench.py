#!/usr/bin/env python
"""bench.py -- SFNO 6 h forecast steps/s at 721x1440x73 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload sfno12_nonlinear|sfno12_linear|filter_linear] [--precision fp32|tf32]

A "step" is one pass of the hot path over one batch of synthetic input: one full 12-block SFNO forward
(FourierNeuralOperatorNet defaults: 73 variables, embed 256, scale_factor 6 -> lmax 120 / mmax 121), batch 1
per GPU, random-init weights.  `value` times it with the input resident in HBM; `e2e` times the same call
through the public nn.Module API with HOST (pinned) input and output buffers, copies inside the timed
region.  N > 1: one process per GPU (torchrun), each rank forecasts its own member -- the path shards over
independent members with no data-path collective ("scaling": "weak").

`--impl reference` times the reference's own CPU implementation of the same path on the host cores: the UNMODIFIED
reference modules (staged under the git-ignored oracle/_ref/ by oracle/build_ref.sh so that they travel to the GPU
box) over the restated torch_harmonics dependency (oracle/th_shim.py).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sfno_6h_forecast_steps_per_sec_721x1440x73"
UNIT = "steps/s"
IMG, NVAR, EMBED, NLAYERS, LMAX, MMAX = (721, 1440), 73, 256, 12, 120, 121
NPOS = 7260  # |{(l, m): l >= m}| = modes carrying data


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region (20 ms period)."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th = index, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception as e:
            self.err = repr(e)
            self.th = None

    def _loop(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), get_reasons(self.h)))
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        if self.th is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable: %s" % getattr(self, "err", "?")]}
        self.stop_flag = True
        self.th.join(timeout=1)
        bits = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                0x80: "hw_power_brake_slowdown"}
        reasons = set()
        for _, r in self.rows:
            for b, n in bits.items():
                if r & b:
                    reasons.add(n)
        sm = [c for c, _ in self.rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------- reference arm
def cpu_reference_step_factory(workload):
    """The reference's own CPU implementation of the path for the workload; returns (full, partial, info).

    kind "reference": the UNMODIFIED reference modules (MSFNO/Models/sfno/{sfnonet,layers,contractions,activations}.py,
    staged byte for byte under oracle/_ref/ by oracle/build_ref.sh, or the mounted /root/reference) driven through their
    public API -- FourierNeuralOperatorNet(device, cfg, filter_type=...).forward(x) -- with the un-vendored
    torch_harmonics dependency restated by oracle/th_shim.py.  kind "port": the functional oracle port
    (oracle/sfno_oracle.py), used only when no copy of the reference is present."""
    import torch
    from oracle import ref_import, sfno_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    if workload == "filter_linear":
        raise SystemExit("--impl reference supports the sfno12_* workloads")
    ftype = "linear" if workload == "sfno12_linear" else "non-linear"
    x = torch.randn(1, NVAR, *IMG, generator=g)
    if ref_import.available():
        ref = ref_import.load()
        ref_import.use_harmonics(ref.th_shim)
        torch.manual_seed(0)
        net = ref.sfnonet.FourierNeuralOperatorNet("cpu", ref.Attributes(), filter_type=ftype).eval()

        def full():
            with torch.no_grad():
                return net(x)

        def partial(nblocks):
            """encoder + first nblocks blocks (bounded sample); timing is extrapolated by the caller."""
            with torch.no_grad():
                h = net.encoder(x) + net.pos_embed
                for blk in list(net.blocks)[:nblocks]:
                    h = blk(h)
                return h

        return full, partial, dict(cores=torch.get_num_threads(), ftype=ftype, kind="reference",
                                   what="unmodified reference modules (%s)" % ref_import.REFERENCE_ROOT)
    tr = sfno_oracle.Transforms(IMG, 6)
    sd = sfno_oracle.make_state_dict(filter_type=ftype, img_size=IMG, in_chans=NVAR, out_chans=NVAR, embed=EMBED,
                                     num_layers=NLAYERS, seed=0)

    def full():
        with torch.no_grad():
            return sfno_oracle.sfno_forward(x, sd, tr, ftype, NLAYERS)

    def partial(nblocks):
        """encoder + first nblocks blocks (bounded sample); timing is extrapolated by the caller."""
        with torch.no_grad():
            h = sfno_oracle.mlp_1x1(x, sd, "encoder.") + sd["pos_embed"]
            for i in range(nblocks):
                h = sfno_oracle.block_forward(h, sd, i, NLAYERS, ftype, tr)
            return h

    return full, partial, dict(cores=torch.get_num_threads(), ftype=ftype, kind="port", what="oracle port (no reference copy present)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    full, partial, info = cpu_reference_step_factory(args.workload)
    t0 = time.perf_counter()
    full()
    t_full = time.perf_counter() - t0  # always one complete forward first (also the warm-up)
    budget = 200.0
    total = args.steps + max(args.warmup - 1, 0)
    if t_full * total <= budget:
        sample, scale, fn = "1 full 12-block forward per step; " + info["what"], 1.0, full
    else:
        # bounded sample: encoder + the first nb blocks; scaled to a full step by the measured ratio
        nb = 2
        t0 = time.perf_counter()
        partial(nb)
        t_part = time.perf_counter() - t0
        scale = t_full / t_part
        sample = "encoder + first %d of 12 blocks per step, scaled x%.2f to a full forward (ratio measured on one full forward); %s" % (nb, scale, info["what"])
        fn = lambda: partial(nb)
    for _ in range(max(args.warmup - 1, 0)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) * scale
    value = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, precision):
    names = {"sfno12_nonlinear": "configs[1]: full SFNO 12-block forward, 73 vars, embed 256, 721x1440, batch 1/GPU, "
                                 "filter_type=non-linear (SpectralAttentionS2, the reference default main.py runs)",
             "sfno12_linear": "configs[1] variant: full SFNO 12-block forward with filter_type=linear (SpectralConvS2, 11.7 G params)",
             "filter_linear": "configs[0]: single spectral filter RealSHT->SpectralConvS2->InverseRealSHT on 721x1440, 256 ch, batch 1"}
    return {"workload": names[args.workload], "precision_tier": precision, "batch_per_gpu": 1, "lmax": LMAX, "mmax": MMAX,
            "l2_policy": "inputs larger than L2 (x 303 MB, activations 1.06 GB per pass vs 126 MB L2)",
            "parallelism": "independent members, one per GPU, no data-path collective"}


def ncu_traffic(csv_names, kernel_substr, which=0):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes) of the `which`-th launch whose kernel name contains
    `kernel_substr`, read from the first of `csv_names` that exists under profiles/ (an `ncu --set full ... --page raw
    --csv` export).  Returns (bytes or None, source or None): a missing capture gives null, never a literal."""
    import csv
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for name in csv_names:
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        try:
            rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
            h = rows[0]
            ik, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            units = rows[1]
            hits = [r for r in rows[2:] if kernel_substr in r[ik]]
            if len(hits) > which:
                r = hits[which]
                val = float(r[ir].replace(",", "")) * unit[units[ir]] + float(r[iw].replace(",", "")) * unit[units[iw]]
                return val, "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of this launch, profiles/%s" % name
        except (ValueError, KeyError, IndexError):
            continue
    return None, None


# ----------------------------------------------------------------------------------------- our arm
def build_ours(args, dev):
    import torch
    import msfno_b200
    msfno_b200.set_precision(args.precision)
    torch.manual_seed(0)
    if args.workload == "filter_linear":
        sht = msfno_b200.RealSHT(*IMG, lmax=LMAX, mmax=MMAX, grid="equiangular").float()
        isht = msfno_b200.InverseRealSHT(*IMG, lmax=LMAX, mmax=MMAX, grid="equiangular").float()
        sht.weights = sht.weights * 1e5
        isht.pct = isht.pct / 1e5
        mod = msfno_b200.SpectralConvS2(sht, isht, EMBED, use_complex_kernels=True)
        mod.forward_transform, mod.inverse_transform = sht, isht
        return mod.to(dev).eval(), (1, EMBED, *IMG)
    ftype = "linear" if args.workload == "sfno12_linear" else "non-linear"
    if ftype == "linear":
        # build directly on the device: 11.7 G parameters (46.8 GB) never touch host memory
        with torch.device(dev):
            net = msfno_b200.FourierNeuralOperatorNet(dev, None, filter_type=ftype)
    else:
        net = msfno_b200.FourierNeuralOperatorNet(dev, None, filter_type=ftype)
    return net.to(dev).eval(), (1, NVAR, *IMG)


def dominant_kernel_roofline(args, net, dev, pk):
    """Roofline of the kernel with the largest share of the step, timed live with CUDA events on the launching
    stream, L2 flushed between launches.  Share evidence: profiles/*launches*.csv."""
    import torch
    from msfno_b200 import _lib
    from msfno_b200._lib import lib, ptr, check
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)

    def timed(fn, iters=8):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    stages = {}
    if args.workload != "filter_linear":
        # BASELINE metric (ii): SHT / ISHT achieved rate vs the HBM roof at the full grid (C = 256, B = 1)
        import msfno_b200
        x = torch.randn(1, EMBED, *IMG, device=dev)
        sht, isht = net.trans_down, net.itrans_up
        with torch.no_grad():
            pm = sht.forward_packed(x)
            cm = msfno_b200.sht.relayout(pm, sht, _lib.LAYOUT_PM, _lib.LAYOUT_CM, 1, EMBED)
            gb = (4.0 * EMBED * IMG[0] * IMG[1] + 8.0 * EMBED * LMAX * MMAX + 4.0 * MMAX * LMAX * IMG[0]) / 1e9
            for name, fn in (("sht_fwd_full", lambda: sht.forward_packed(x)), ("isht_fwd_full", lambda: isht.inverse_packed(cm))):
                ms = timed(fn)
                stages[name] = {"ms": ms, "algorithmic_GB": gb, "achieved_GBps": gb / ms * 1e3, "frac_of_hbm_peak": gb / ms * 1e3 / pk["hbm"],
                                "launches": 2}
        del x, pm, cm
    if args.workload == "sfno12_nonlinear":
        # spectral complex-MLP hidden layer: real GEMM [P x 1024] x [1024 x 1024] (2 of the 4 MLP launches per block)
        Pp = 7440
        A = torch.randn(Pp, 1024, device=dev)
        W = torch.randn(1024, 1024, device=dev)
        D = torch.empty(Pp, 1024, device=dev)
        prec = _lib.PREC_TF32 if args.precision == "tf32" else _lib.PREC_FP32
        ms = timed(lambda: check(lib.msfno_gemm_nt(ptr(A), 1024, ptr(W), 1024, ptr(D), 1024, Pp, 1024, 1024, 1, prec, st)))
        flops = 2.0 * NPOS * 1024 * 1024  # algorithmic: only the 7260 modes with l >= m
        ach = flops / (ms * 1e-3) / 1e12
        peak = pk["bf16_sustained"] / 2.0
        stages["spectral_mlp_hidden_gemm"] = {
            "kernel": "%s: [7440 x 1024] x [1024 x 1024], 24 launches per step" % (
                "gemm_tc2_kernel (CTA pair, cta_group::2)" if args.precision == "tf32" else "gemm_tc3_kernel (3xTF32, register accumulation)"), "ms": ms,
            "bound": "tensor", "achieved_TFLOPs": ach, "peak_TFLOPs": peak, "frac_of_tf32_peak": ach / peak,
            "peak_source": "%s bf16_tflops_sustained / 2 (TF32 = half of BF16, BASELINE.md section 2)" % pk["src"]}
        del A, W, D
        if args.precision == "tf32" and hasattr(net, "_encode_fused"):
            # dominant kernel by share of the step (profiles/r01_launches_bench_tf32_v7.csv: mlp_tc_kernel 31 %): its largest
            # launch, the fused encoder MLP  y = W2 gelu(W1 x + b1) + b2 + pos_embed  at 721x1440
            xin = torch.randn(1, net.in_chans, *IMG, device=dev)
            with torch.no_grad():
                ms = timed(lambda: net._encode_fused(xin))
            by = 4.0 * IMG[0] * IMG[1] * (net.in_chans + 2 * EMBED)   # input + pos_embed read, output written, once each
            ach = by / (ms * 1e-3) / 1e9
            traffic, tsrc = ncu_traffic(["r02_ncu_mlp_tc_raw.csv", "r01_ncu_final_raw.csv"], "mlp_tc_kernel", 0)
            return {"kernel": "mlp_tc_kernel (fused encoder MLP 73->256->256 + pos_embed, hidden tile in TMEM; 1 of 13 launches per step)",
                    "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": traffic,
                    "traffic_source": tsrc,
                    "peak_source": "%s hbm_gbs" % pk["src"], "ms_per_launch": ms, "algorithmic_bytes": by, "stages": stages}
        traffic, tsrc = ncu_traffic(["r02_ncu_x3_gemm_raw.csv"], "gemm_tc3_kernel", 0)
        return {"kernel": "gemm_tc3_kernel (3xTF32 spectral complex-MLP hidden layer, M=7260 modes, N=K=1024 real)", "bound": "tensor",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": tsrc,
                "peak_source": "%s bf16_tflops_sustained / 2 (TF32 = half of BF16, BASELINE.md section 2)" % pk["src"],
                "ms_per_launch": ms, "stages": stages}
    # linear workloads: the SpectralConvS2 weight stream
    import msfno_b200
    filt = net if args.workload == "filter_linear" else net.blocks[1].filter_layer.filter
    sht = filt.forward_transform
    plan = sht._get_plan(dev)
    a_pm = torch.randn(1, plan.P, 2 * EMBED, device=dev)
    out = torch.empty(1, plan.P, 2 * EMBED, device=dev)
    w = filt.w.detach()
    ws = torch.empty(lib.msfno_specconv_ws_floats(plan.h, 1, EMBED, EMBED), device=dev)
    ms = timed(lambda: check(lib.msfno_specconv_fwd(plan.h, ptr(a_pm), ptr(w), ptr(out), ptr(ws), 1, EMBED, EMBED, st)))
    by = 8.0 * EMBED * EMBED * NPOS + 16.0 * EMBED * NPOS
    ach = by / (ms * 1e-3) / 1e9
    return {"kernel": "specconv_gather_kernel + specconv_tma_kernel (per-mode complex channel contraction, 3.8 GB weight stream)", "bound": "hbm",
            "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": None,
            "peak_source": "%s hbm_gbs (a COPY figure: half reads, half writes; this kernel is a pure read stream, which "
                           "HBM3e serves slightly faster, so frac can exceed 1)" % pk["src"],
            "ms_per_launch": ms, "stages": stages}


def copy_ceiling(dev, nbytes, iters=5):
    """Bare pinned-memory copies of the e2e payload on this rank: H2D and D2H of `nbytes` each on two streams at once
    (what HostPipeline overlaps), no kernels.  Returns ms per (H2D + D2H) pair: e2e cannot beat it."""
    import torch
    h_in = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    h_out = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    d_in = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    d_out = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    for rep in range(2):   # first pass warms the pinned mappings
        e0.record(main)
        s1.wait_stream(main)
        s2.wait_stream(main)
        for _ in range(iters):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        main.wait_stream(s1)
        main.wait_stream(s2)
        e1.record(main)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the MSFNO hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import msfno_b200
    from msfno_b200 import _lib
    from msfno_b200.pipeline import bind_host_to_device
    # before any pinned allocation: this rank's CPUs (and first-touch pages) = the NUMA node of its GPU
    cpu_binding = bind_host_to_device(dev) if world > 1 else None
    pk = peaks()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g = torch.Generator().manual_seed(1234 + rank)
    xshape = (1, EMBED, *IMG) if args.workload == "filter_linear" else (1, NVAR, *IMG)
    x_host = torch.randn(*xshape, generator=g).pin_memory()
    x_dev = x_host.to(dev)

    def measure(precision, want_clocks):
        """One tier: device-resident leg (value) and host-buffer leg (e2e), each K steps after W warm-ups."""
        targs = argparse.Namespace(**vars(args))
        targs.precision = precision
        net, _ = build_ours(targs, dev)
        with torch.no_grad():
            y = net(x_dev)
            y_host = torch.empty(y.shape, dtype=y.dtype).pin_memory()
            eager = net
            launches_per_step = None
            if not args.no_graph:
                l_one = _lib.lib.msfno_launch_count()
                eager(x_dev)
                launches_per_step = _lib.lib.msfno_launch_count() - l_one   # kernels one replay contains
                net = msfno_b200.GraphedForward(eager, x_dev)
            for _ in range(max(args.warmup - 1, 0)):
                net(x_dev)
            # ---- device-resident leg ---------------------------------------------------------------
            sampler = ClockSampler(local)
            sync_all()
            if rank == 0 and want_clocks:
                sampler.start()
            l0 = _lib.lib.msfno_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if args.no_graph:
                for _ in range(args.steps):
                    y = net(x_dev)
            else:  # the input already sits in the graph's static input buffer (resident in HBM)
                for _ in range(args.steps):
                    y = net()
            e1.record()
            sync_all()
            launches = _lib.lib.msfno_launch_count() - l0
            if not args.no_graph:
                launches = launches_per_step * args.steps  # replayed inside the graph: the library counter does not see replays
            ms_dev = max_over_ranks(e0.elapsed_time(e1))
            clocks = sampler.stop() if (rank == 0 and want_clocks) else None
            # ---- end-to-end leg: host buffers in, host buffers out, through the public HostPipeline API ----------
            pipe = msfno_b200.HostPipeline(net, dev)
            xs = [x_host, x_host.clone().pin_memory()]
            ys = [y_host, torch.empty_like(y_host).pin_memory()]
            pipe.run([xs[i % 2] for i in range(3)], [ys[i % 2] for i in range(3)])
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pipe.run([xs[i % 2] for i in range(args.steps)], [ys[i % 2] for i in range(args.steps)])
            e1.record()
            sync_all()
            ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        return dict(net=net, eager=eager, ms_dev=ms_dev, ms_e2e=ms_e2e, launches=int(launches), clocks=clocks,
                    finite=bool(torch.isfinite(y_host).all()), h2d=x_host.numel() * 4, d2h=y_host.numel() * 4)

    primary = measure(args.precision, True)
    with torch.no_grad():
        roof = dominant_kernel_roofline(args, primary["eager"], dev, pk) if rank == 0 else None
    ceiling_ms = max_over_ranks(copy_ceiling(dev, primary["h2d"]))
    other_name = "fp32" if args.precision == "tf32" else "tf32"
    other = None
    if not args.one_tier and args.workload == "sfno12_nonlinear":
        del primary["net"], primary["eager"]
        torch.cuda.empty_cache()
        other = measure(other_name, False)
        del other["net"], other["eager"]
        torch.cuda.empty_cache()

    # the configs that shard / communicate (config 5 sharded transform, config 3 DDP step, config 4 ensemble rollout): timed
    # in the same run at every N -- at N = 1 too, so that a 1 -> 8 scaling run carries its own single-GPU baselines
    multi = None
    if not args.no_multi_gpu_extras and args.workload == "sfno12_nonlinear":
        multi = multi_gpu_extras(args, dev, rank, world, max_over_ranks, sync_all)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and args.workload != "filter_linear":
        try:
            full, partial, info = cpu_reference_step_factory(args.workload)
            full()  # warm-up (thread pools, page faults)
            t0 = time.perf_counter()
            full()
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": 1.0 / dt, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                            "sample": "1 full 12-block forward (after 1 warm-up), fp32; " + info["what"]}
        except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU measurement
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}

    def tier_entry(m, name):
        return {"dtype": "f32 (3xTF32 on tcgen05, fp32 accumulate, FFT longitude transforms)" if name == "fp32" else "tf32",
                "tolerance_rel_l2": 1e-5 if name == "fp32" else 2e-3,
                "value": world * args.steps / (m["ms_dev"] * 1e-3), "ms_per_step": m["ms_dev"] / args.steps,
                "e2e": world * args.steps / (m["ms_e2e"] * 1e-3), "e2e_ms_per_step": m["ms_e2e"] / args.steps,
                "gpu_launches": m["launches"], "output_finite": m["finite"]}

    tiers = {args.precision: tier_entry(primary, args.precision)}
    if other is not None:
        tiers[other_name] = tier_entry(other, other_name)
    value = world * args.steps / (primary["ms_dev"] * 1e-3)
    e2e_ms = primary["ms_e2e"] / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": primary["ms_dev"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "tf32", "data": "synthetic",
        "config": workload_config(args, args.precision),
        "e2e": {"value": world * args.steps / (primary["ms_e2e"] * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": primary["h2d"], "d2h_bytes_per_step": primary["d2h"],
                "ms_per_step": e2e_ms,
                # bare pinned H2D + D2H of the same payload on this box at this N (max over ranks), no kernels: the
                # host-link floor of one step; e2e cannot beat it
                "copy_ceiling_ms_per_step": ceiling_ms,
                "copy_ceiling_GBps": world * (primary["h2d"] + primary["d2h"]) / (ceiling_ms * 1e-3) / 1e9,
                "frac_of_copy_ceiling": ceiling_ms / e2e_ms, "rank0_cpu_binding": cpu_binding},
        "tiers": tiers,
        "gpu_launches": primary["launches"],
        "clocks": primary["clocks"],
        "roofline": roof,
        "cpu_baseline": cpu_baseline,
        "output_finite": primary["finite"] and (other is None or other["finite"]),
        "cuda_graph": not args.no_graph,
    }
    if multi is not None:
        line["multi_gpu"] = multi
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def multi_gpu_extras(args, dev, rank, world, max_over_ranks, sync_all):
    """The paths that shard or communicate, timed in the same run at every N (device time, max over ranks).
    sharded_sht: BASELINE configs[4] (A): 1441 x 2880, 256 channels, lmax 240 -- latitude-sharded FFT, all-to-all
                 lat<->m over NVLink, order-sharded Legendre, and back.
    ddp_train:   BASELINE configs[2]: MSFNO fwd + bwd + Adam on the FiLM head, per-rank batch 8, film_layers 1,
                 DistributedDataParallel over NCCL.
    ensemble_rollout: BASELINE configs[3]: 28-day (112-step) autoregressive rollout, 8 members per GPU (64 on 8 GPUs),
                 members sharded across ranks, no data-path collective."""
    import torch
    import torch.distributed as dist
    out = {}
    own_group = False
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_sharded_sht
        if world == 1 and not dist.is_initialized():
            # the sharded transform talks to torch.distributed even with one rank: a private single-rank group
            import socket
            s = socket.socket()
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
            s.close()
            dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1, device_id=dev)
            own_group = True
        out["sharded_sht"] = bench_sharded_sht.run(dev, rank, world, max_over_ranks, sync_all, steps=max(3, min(args.steps, 10)))
    except Exception as e:
        out["sharded_sht"] = {"error": repr(e)}
    finally:
        if own_group:
            try:
                dist.destroy_process_group()
            except Exception:
                pass
    torch.cuda.empty_cache()
    try:
        import bench_train_step
        out["ddp_train"] = bench_train_step.run(dev, rank, world, max_over_ranks, sync_all, batch=8, film_layers=1,
                                                steps=max(2, min(args.steps, 5)))
    except Exception as e:
        out["ddp_train"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    try:
        import bench_rollout
        out["ensemble_rollout"] = bench_rollout.run(dev, rank, world, max_over_ranks, sync_all, members=8, batch=1, steps=112)
    except Exception as e:
        out["ensemble_rollout"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    return out


def main():
    # stdout carries the one JSON line only: native libraries (NCCL prints its version banner there) write to file
    # descriptor 1, so descriptor 1 is pointed at stderr and Python's sys.stdout keeps the original stream
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(keep, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sfno12_nonlinear", choices=["sfno12_nonlinear", "sfno12_linear", "filter_linear"])
    ap.add_argument("--precision", default=os.environ.get("MSFNO_PRECISION", "tf32"), choices=["fp32", "tf32"],
                    help="tf32: tensor-core tier (<= 2e-3 rel-L2 vs the reference, tests/test_gpu_tc.py); fp32: exact tier (<= 1e-5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--one-tier", action="store_true", help="measure only --precision (default: both tiers, reported under `tiers`)")
    ap.add_argument("--no-multi-gpu-extras", action="store_true", help="N > 1: skip the sharded-SHT and DDP-training legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
