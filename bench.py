#!/usr/bin/env python
"""bench.py — samples/sec of one DeepFM Criteo-shape training step on B200 (BASELINE.json
configs[1]): 39 categorical fields x 25,641 rows (~1.0 M rows), emb_dim 16, MLP 300-300-300,
batch 4096 per GPU.  A "step" is one pass of the hot path over one synthetic batch:
zero_grad -> fused gather + LR + FM + MLP forward -> BCE -> backward (dense-gradient
scatter-add, GEMM dgrad/wgrad) -> clip_grad_norm_(10) -> Adam  (the reference's
BaseModel.train_step, fuxictr/pytorch/models/rank_model.py:307-323).

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference ...                   # the reference's CPU path (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from collections import OrderedDict

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NF, VOCAB, DIM, HIDDEN, BATCH = 39, 25641, 16, [300, 300, 300], 4096
METRIC = "samples/sec DeepFM Criteo-shape train_step (fwd+bwd+clip+Adam)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--graph", type=int, default=1, help="capture the step in a CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nbatches", type=int, default=8, help="distinct synthetic batches rotated through")
    ap.add_argument("--dp-only", action="store_true",
                    help="N>1: replicate the tables and all-reduce the whole gradient arena instead of row-sharding")
    ap.add_argument("--lazy-adam", type=int, default=0,
                    help="1: evaluate the dense Adam semantics of the tables row-wise and lazily (bit-identical)")
    ap.add_argument("--steps-only", action="store_true", help="skip the per-kernel / stress / CPU legs (profiling)")
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32x3", "tf32"],
                    help="arithmetic of the dense GEMMs: fp32 FFMA, 3xTF32 (fp32-class) or 1xTF32 on tcgen05")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "DeepFM Criteo-shape: %d fields x %d rows, emb_dim %d, MLP %s, fp32 Adam"
                        % (NF, VOCAB, DIM, HIDDEN),
            "global_batch": args.batch * n_gpus, "per_gpu_batch": args.batch,
            "parallelism": ("single" if n_gpus == 1 else
                            ("dp%d (replicated tables, all-reduce of the gradient arena)" % n_gpus if args.dp_only else
                             "tables row-sharded over %d GPUs (P2P push/pull over NVLink) + dense dp all-reduce" % n_gpus)),
            "cache": "working set (4 fp32 arenas x 68 MB: params, grads, Adam m/v) exceeds the 126 MB L2; "
                     "%d distinct index batches are rotated" % args.nbatches}


def make_specs():
    return [("C%d" % i, {"type": "categorical", "source": "", "padding_idx": 0, "vocab_size": VOCAB})
            for i in range(NF)]


def make_batches(n, batch, seed=0):
    """uniform ids in [1, V), Bernoulli(0.25) labels, one (B, F+1) float64 matrix
    per batch — exactly what the reference's BatchCollator hands to the model."""
    import torch
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        ids = torch.randint(1, VOCAB, (batch, NF), generator=gen).double()
        label = (torch.rand(batch, 1, generator=gen) < 0.25).double()
        out.append(torch.cat([ids, label], dim=1))
    return out


# ------------------------------------------------------------------------------------------
# clocks / throttle sampling during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows, self.proc, self.dev = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU leg: the reference's own path (ATen ops on host cores), via the oracle restatement
# ------------------------------------------------------------------------------------------
def cpu_reference_run(batch, steps, warmup, seconds=None, threads=None):
    """Times BaseModel.train_step as the reference executes it on CPU (oracle port: the same ATen
    ops — F.embedding x39, stack, FM, LR, Linear/ReLU, BCE, autograd, clip_grad_norm_, Adam)."""
    import torch
    from oracle import fuxictr_oracle as O
    from fuxictr_b200 import zoo
    from fuxictr_b200.schema import FeatureMap
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    specs = make_specs()
    fm = FeatureMap.from_specs(specs, embedding_dim=DIM)
    torch.manual_seed(2019)
    model = zoo.DeepFM(fm, gpu=-1, embedding_dim=DIM, hidden_units=HIDDEN)  # parameter container only
    spec_map = OrderedDict(specs)
    tr = O.OracleTrainer(model.state_dict(), lambda s, X: torch.sigmoid(O.deepfm_logit(spec_map, s, X, len(HIDDEN))),
                         spec_map, ["label"])
    batches = [fm.batch_dict(m) for m in make_batches(4, batch)]
    if threads is None:
        # "all the host threads it can use": ATen's intra-op pool stops scaling (and then collapses)
        # long before 100+ threads on ops this small, so give the reference the thread count at
        # which IT runs fastest on this box, and report that count.
        best = (0.0, 1)
        for cand in sorted(set(c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu)):
            torch.set_num_threads(cand)
            tr.train_step(batches[0])
            t0 = time.perf_counter()
            tr.train_step(batches[1])
            rate = 1.0 / (time.perf_counter() - t0)
            if rate > best[0]:
                best = (rate, cand)
        threads = best[1]
    torch.set_num_threads(threads)
    for i in range(warmup):
        tr.train_step(batches[i % len(batches)])
    t0 = time.perf_counter()
    done = 0
    while True:
        tr.train_step(batches[done % len(batches)])
        done += 1
        el = time.perf_counter() - t0
        if seconds is not None:
            if el >= seconds and done >= 3:
                break
        elif done >= steps:
            break
    return {"value": batch * done / el, "ms_per_step": 1e3 * el / done, "steps": done, "cores": threads, "host_cpus": ncpu}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded: each step is one 4096-sample batch; cap the run at a few minutes
    steps = min(args.steps, 400)
    r = cpu_reference_run(args.batch, steps, min(args.warmup, 5))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": r["steps"], "warmup": min(args.warmup, 5),
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, 1),
            "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                             "sample": "%d train steps of batch %d on %d of %d host threads (fastest setting; oracle "
                                       "port of the reference's ATen path)" % (r["steps"], args.batch, r["cores"],
                                                                               r["host_cpus"])},
            "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def time_kernel(fn, reps, stream_sync=None):
    """Average device duration (ms) of `fn` over `reps` back-to-back launches.  The launches are
    captured into one CUDA graph so that host-side (Python/ctypes) launch cost does not leak into
    a microsecond-scale kernel time; CUDA events bracket the replay on the launching stream."""
    import torch
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def kernel_rooflines(model, fm, dev_batch, peaks, args, with_gather=True):
    """Per-kernel achieved bandwidth, measured live with CUDA events: the fused gather (north-star
    kernel) and the dense clip+Adam pass (largest share of the step)."""
    import ctypes
    import torch
    from fuxictr_b200 import _lib, functional as F2
    out = {}
    hbm = peaks["hbm_gbs"]
    X = OrderedDict((k, v) for k, v in fm.batch_dict(dev_batch).items() if k != "label")
    B = dev_batch.shape[0]
    # fused multi-field gather, algorithmic bytes: F*8 (ids) + F*D*4 (rows) + F*D*4 (out)
    fed = model.embedding_layer.embedding_layer
    if with_gather:
        with torch.no_grad():
            ms = time_kernel(lambda: fed.forward_tensor(X), 50, None)
        gbytes = B * (NF * 8 + 2 * NF * DIM * 4)
        out["embed_gather_fwd"] = {"ms": ms, "algorithmic_bytes": gbytes, "GBps": gbytes / ms / 1e6,
                                   "frac_of_measured_hbm": gbytes / ms / 1e6 / hbm,
                                   "note": "B=%d: tables (64 MB) are L2-resident and the launch is latency-bound" % B}
    # dense optimizer pass over the arena: 4 reads (p,g,m,v) + 4 writes (p,m,v, g=0) of fp32
    opt = model._fused_optimizer
    a = model._arena
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    P, G, M, V = (ctypes.c_void_p(t.data_ptr()) for t in (a.P, a.G, opt.M, opt.V))
    backup = (a.P.clone(), opt.M.clone(), opt.V.clone(), a.G.clone())
    opt.step_dev.add_(1)
    ms = time_kernel(lambda: _lib.call("b2_adam_step", P, G, M, V, a.numel, ctypes.c_void_p(opt.sumsq.data_ptr()),
                                       10.0, 1e-3, 0.9, 0.999, 1e-8, ctypes.c_void_p(opt.step_dev.data_ptr()),
                                       1, st()), 30, None)
    abytes = a.numel * 4 * 8
    out["adam_step"] = {"ms": ms, "algorithmic_bytes": abytes, "GBps": abytes / ms / 1e6,
                        "frac_of_measured_hbm": abytes / ms / 1e6 / hbm}
    ms = time_kernel(lambda: _lib.call("b2_sumsq", G, a.numel, ctypes.c_void_p(opt.sumsq.data_ptr()), st()), 30, None)
    sbytes = a.numel * 4
    out["grad_sumsq"] = {"ms": ms, "algorithmic_bytes": sbytes, "GBps": sbytes / ms / 1e6,
                         "frac_of_measured_hbm": sbytes / ms / 1e6 / hbm}
    a.P.copy_(backup[0]); opt.M.copy_(backup[1]); opt.V.copy_(backup[2]); a.G.copy_(backup[3])
    opt.step_dev.sub_(1)
    # the tensor-core GEMM of the first MLP layer (4096 x 300 x 624), the most frequent kernel of the step
    mode = F2.get_matmul_precision()
    if mode != "fp32":
        M_, N_, K_ = B, HIDDEN[0], NF * DIM
        xa = torch.randn(M_, K_, device="cuda"); wb = torch.randn(N_, K_, device="cuda")
        yo = torch.empty(M_, N_, device="cuda")
        xs = F2.split_tf32(xa) if mode == "tf32x3" else None
        ws = F2.split_tf32(wb) if mode == "tf32x3" else None
        ms = time_kernel(lambda: F2.gemm_nt(xa, wb, yo, a_small=xs, b_small=ws), 30, None)
        flops = 2.0 * M_ * N_ * K_ * (3 if mode == "tf32x3" else 1)
        tf32_peak = peaks.get("bf16_tflops", 1590.0) / 2.0       # dense TF32 = half the measured bf16 rate
        out["gemm_tf32"] = {"ms": ms, "shape": [M_, N_, K_], "passes": 3 if mode == "tf32x3" else 1,
                            "TFLOPs": flops / ms / 1e9, "frac_of_tf32_peak": flops / ms / 1e9 / tf32_peak,
                            "note": "launch + L2->SM operand traffic bound at this size (128 CTAs, 20 k-blocks each)"}
    return out


def gather_stress(peaks):
    """The gather against a table set far larger than L2 (39 x 4 M rows x 64 B = 10 GB) at large
    batch: the HBM-bound operating point BASELINE.md asks the 60 %-of-peak claim to be made on."""
    import torch
    from fuxictr_b200 import functional as F2
    res = []
    rows, F_, D = 4_000_000, NF, DIM
    tables = [torch.empty(rows, D, device="cuda").normal_(0, 0.01) for _ in range(F_)]
    plan = F2.GatherPlan([F2.GatherField("C%d" % i, i, D, padding_idx=0) for i in range(F_)])
    for B in (4096, 65536, 524288):
        gen = torch.Generator(device="cuda").manual_seed(B)
        mat = torch.randint(1, rows, (B, F_ + 1), device="cuda", generator=gen).double()
        idx = [mat[:, i] for i in range(F_)]
        with torch.no_grad():
            ms = time_kernel(lambda: F2.embed_gather(plan, idx, tables), 20, None)
        nbytes = B * (F_ * 8 + 2 * F_ * D * 4)
        res.append({"batch": B, "ms": ms, "GBps": nbytes / ms / 1e6,
                    "frac_of_measured_hbm": nbytes / ms / 1e6 / peaks["hbm_gbs"]})
        del mat, idx
    del tables
    torch.cuda.empty_cache()
    return {"table_bytes": rows * D * 4 * F_, "points": res}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fd:
            p = json.load(fd)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops", 1590.0)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from fuxictr_b200 import zoo, _lib, functional as F2
    from fuxictr_b200.schema import FeatureMap
    _lib.load()
    F2.set_matmul_precision(args.precision)
    peaks = load_peaks()

    fm = FeatureMap.from_specs(make_specs(), embedding_dim=DIM)
    torch.manual_seed(2019)
    model = zoo.DeepFM(fm, gpu=local, embedding_dim=DIM, hidden_units=HIDDEN)
    sharded = world > 1 and not args.dp_only
    if sharded:
        try:
            from fuxictr_b200.sharded import SymmPeerGroup
            model.enable_sharding(SymmPeerGroup(), args.batch, NF + 1, torch.float64)
        except Exception as exc:   # no peer-mapped memory on this box: replicated tables + arena all-reduce
            sys.stderr.write("[bench r%d] row-sharding unavailable (%r); using --dp-only\n" % (rank, exc))
            sharded = False
            args.dp_only = True
            torch.manual_seed(2019)
            model = zoo.DeepFM(fm, gpu=local, embedding_dim=DIM, hidden_units=HIDDEN)
    opt = model.use_fused_optimizer(lazy_tables=bool(args.lazy_adam) and world == 1)
    if world > 1 and not sharded:
        opt.grad_allreduce = True
    model.train()
    host_batches = [m.pin_memory() for m in make_batches(args.nbatches, args.batch, seed=1000 + rank)]
    dev_batches = [m.cuda(non_blocking=True) for m in host_batches]
    # launches per step (our kernels only): counted by wrapping the C-ABI call around one eager step
    from fuxictr_b200.pipeline import TrainPipeline
    pipe = TrainPipeline(model, args.batch, NF + 1, torch.float64, graph=False)
    pipe.prime(dev_batches[0])
    counter = {"n": 0}
    orig_call = _lib.call

    def counting_call(name, *a):
        counter["n"] += 1
        return orig_call(name, *a)

    if os.environ.get("B2_BENCH_VERBOSE"):
        sys.stderr.write("[bench r%d] model built\n" % rank); sys.stderr.flush()
    _lib.call = counting_call
    pipe.step_device(dev_batches[0])
    torch.cuda.synchronize()
    launches = counter["n"]
    _lib.call = orig_call
    if args.graph:
        try:
            pipe.capture(3)     # the constructor does this for graph=True; done late here to count launches first
        except Exception as exc:   # report an eager number rather than none (the JSON line says cuda_graph: false)
            sys.stderr.write("[bench r%d] CUDA-graph capture failed (%r); timing the eager step\n" % (rank, exc))
            pipe.graph, pipe.loss_dev = None, None
            torch.cuda.synchronize()

    if os.environ.get("B2_BENCH_VERBOSE"):
        sys.stderr.write("[bench r%d] graph captured\n" % rank); sys.stderr.flush()

    def run_step(i, e2e):
        if e2e:     # the user-facing call: pinned host matrix -> async H2D (overlapping the previous
            pipe.step(host_batches[i % len(host_batches)])     # step) -> graph replay -> loss D2H
        else:       # inputs already resident in HBM
            pipe.step_device(dev_batches[i % len(dev_batches)])

    def timed(e2e, steps, warmup):
        for i in range(warmup):
            run_step(i, e2e)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            run_step(warmup + i, e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    warmup = max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(False, args.steps, warmup)
    ms_e2e = timed(True, args.steps, warmup)
    clocks = sampler.stop() if rank == 0 else None
    final_loss = pipe.loss()

    def note(msg):
        if os.environ.get("B2_BENCH_VERBOSE"):
            sys.stderr.write("[bench r%d] %s\n" % (rank, msg))
            sys.stderr.flush()
    note("timed regions done")
    if rank != 0:
        # Last collective done.  Leave without NCCL teardown: rank 0 still captures CUDA graphs for
        # its single-rank kernel measurements, and a concurrent communicator finalize hung here.
        sys.stdout.flush()
        os._exit(0)
    samples = args.batch * world * args.steps
    value = samples / (ms_total / 1e3)
    e2e_value = samples / (ms_e2e / 1e3)
    if args.steps_only:
        print(json.dumps({"value": value, "ms_per_step": ms_total / args.steps, "e2e": e2e_value,
                          "gpu_launches_per_step": launches, "precision": args.precision, "n_gpus": world}))
        sys.stdout.flush()
        if world > 1:
            os._exit(0)
        return
    kernels = kernel_rooflines(model, fm, dev_batches[0], peaks, args, with_gather=not sharded)
    dom = max([k for k in ("adam_step", "embed_gather_fwd", "grad_sumsq") if k in kernels],
              key=lambda k: kernels[k]["ms"])
    step_ms = ms_total / args.steps
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_step_kernels_traffic.json")
    if os.path.exists(tpath):   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture
        with open(tpath) as fd:
            cap = json.load(fd).get({"adam_step": "adam_kernel", "grad_sumsq": "sumsq_kernel"}.get(dom, dom))
        if cap:
            traffic = (cap["dram_read_MB"] + cap["dram_write_MB"]) * 1e6
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["GBps"], "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": kernels[dom]["frac_of_measured_hbm"], "traffic": traffic,
                "algorithmic_bytes": kernels[dom]["algorithmic_bytes"],
                "peak_source": peaks["source"], "share_of_step": kernels[dom]["ms"] / step_ms}
    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32x3": "f32 (3xTF32 tensor-core GEMMs, fp32 everything else)",
                      "tf32": "tf32 GEMMs, f32 elsewhere"}[args.precision],
            "data": "synthetic", "config": dict(workload_config(args, world), matmul=args.precision,
                                             adam=("dense semantics, lazy row-wise evaluation" if args.lazy_adam and world == 1
                                                   else "dense pass over the arena")),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": pipe.h2d_bytes_per_step * world, "d2h_bytes_per_step": pipe.d2h_bytes_per_step * world,
                    "api": "fuxictr_b200.pipeline.TrainPipeline.step (double-buffered H2D on a copy stream)"},
            "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
            "cuda_graph": pipe.graph is not None, "final_loss": final_loss,
            "roofline": roofline, "kernels": kernels}
    if world == 1:
        try:
            line["gather_stress"] = gather_stress(peaks)
        except Exception as exc:  # e.g. a smaller GPU: report, do not hide
            line["gather_stress"] = {"error": str(exc)}
        if not args.no_cpu_baseline:
            r = cpu_reference_run(args.batch, 0, 3, seconds=args.cpu_seconds)
            line["cpu_baseline"] = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                                    "sample": "%d train steps of batch %d in %.0f s on %d of %d host threads (fastest "
                                              "setting; oracle port of the reference's ATen path)"
                                              % (r["steps"], args.batch, args.cpu_seconds, r["cores"], r["host_cpus"])}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)
