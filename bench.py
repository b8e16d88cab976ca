#!/usr/bin/env python
"""bench.py — ParSeNet hot path on B200:  python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path (BASELINE.json metric) over one batch of synthetic shapes per GPU:
segmentation network forward (3 kNN graphs + edge-convs + head), triplet + NLL losses, Evaluation.fitting_loss
(mean-shift clustering, Hungarian match, primitive / spline fit, residual), backward, gradient all-reduce (N>1), Adam.
Prints ONE JSON line (rank 0).  `value` = shapes/s with inputs resident in HBM; `e2e` = same through the public
module API with pinned-host inputs copied H2D and the loss read back D2H inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "parsenet-codebase_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_POINTS = 10000
BATCH_PER_GPU = 16
KNN_K = 80
EMB = 128
N_PRIM = 10
N_PATCHES = 8          # per shape: plane, sphere, cylinder, cone, open spline, closed spline, plane, sphere


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled every 100 ms while the timed loops run.  In-process NVML (pynvml:
    initialised once, before the warm-up; a query is two cheap driver calls) -- the `nvidia-smi -lms 100` child process used
    before stalled kernel launches for 0.5 - 1 s some seconds after its start (NVML initialisation and enumeration of every GPU
    of the box in a second process), which inflated the first timed steps on some boxes (profiles/r02_scaling.md).  Falls back
    to nvidia-smi when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []          # (time, sm_mhz, sm_max_mhz, set of reasons)
        self.proc = None
        self.idx = gpu_index
        self.nvml = None
        self.stop_flag = False
        self.t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES re-numbers the devices torch sees: resolve the physical device through its UUID-free PCI id
            try:
                bus = torch.cuda.get_device_properties(self.idx).pci_bus_id
                dom = torch.cuda.get_device_properties(self.idx).pci_domain_id
                dev = torch.cuda.get_device_properties(self.idx).pci_device_id
                h = pynvml.nvmlDeviceGetHandleByPciBusId("%08x:%02x:%02x.0" % (dom, bus, dev))
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nvml = (pynvml, h)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nvml
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
        bits = [(n, getattr(pynvml, a, None) or getattr(pynvml, b_, 0)) for n, a, b_ in names]
        while not self.stop_flag:
            try:
                sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((time.time(), sm, self.smax, {n for n, bit in bits if bit and (mask & bit)}))
            except Exception:
                pass
            time.sleep(0.1)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            try:
                sm, smax = float(f[1]), float(f[2])
            except Exception:
                continue
            reasons = {n for n, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9])
                       if v.lower().startswith("active")}
            self.rows.append((time.time(), sm, smax, reasons))

    def stop(self, t0, t1):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, v, vmax, rs in self.rows:
            if ts < t0 or ts > t1:
                continue
            sm.append(v); smax = vmax
            reasons |= rs
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml (in-process)" if self.nvml is not None else "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------ workload
def make_host_batch(B, N, seed):
    """raw host batch as the reference's loader holds it BEFORE its per-batch preprocessing: points, normals (B,N,3), labels,
    primitive ids (B,N), in pinned memory"""
    from tools.synth import ALL_KINDS, synth_cloud   # plain numpy generator (not oracle code)
    pts, nrm, lab, prim = synth_cloud(B, N, seed=seed, n_patches=N_PATCHES, kinds=ALL_KINDS)
    return (torch.from_numpy(pts).pin_memory(), torch.from_numpy(nrm).pin_memory(), torch.from_numpy(lab).pin_memory(),
            torch.from_numpy(prim).pin_memory())


def device_input(pts_d, nrm_d, R_d):
    """the per-batch preprocessing of Dataset.get_train (align the minor principal axis with x, scale to unit extent;
    pnb200/input_pipeline.py) on the device, then the (B,6,N) layout the network is fed"""
    from pnb200.input_pipeline import preprocess_on_device
    p, n = preprocess_on_device(pts_d, nrm_d, R_d)
    return torch.cat([p, n], 2).permute(0, 2, 1).contiguous()


PIN_ALPHA = 3.0


def cluster_codes(device=None):
    """fixed unit codes, one per ground-truth patch id (seeded; the same on every rank and in the CPU arm)"""
    g = torch.Generator().manual_seed(1234)
    c = torch.nn.functional.normalize(torch.randn(64, EMB, generator=g), dim=1)
    return c if device is None else c.to(device)


def pin_clusters(emb_bdn, lab_bn, codes):
    """Workload pin: with random-init weights the embedding clusters into 1-8 segments per shape
    depending on the Adam step, so the fit stage (the launch-bound part) was under-weighted and drifting.  Adding a fixed
    code of the point's ground-truth patch, PIN_ALPHA x the embedding's own RMS norm, makes mean-shift recover exactly
    the N_PATCHES patches of every shape at every step (checked on the CPU port for alpha >= 2).  The addition is outside
    the product (the product is handed an embedding either way) and the gradient flows through it unchanged."""
    scale = emb_bdn.detach().pow(2).mean().sqrt() * (PIN_ALPHA * EMB ** 0.5)
    return emb_bdn + scale * codes[lab_bn].permute(0, 2, 1)


def seeded_splinenet(mode, seed, device):
    """SplineNet with seeded random weights (the pretrained open/closed_spline.pth files are not available)"""
    from src.model import DGCNNControlPoints
    torch.manual_seed(seed)
    return DGCNNControlPoints(20, num_points=10, mode=mode).to(device).eval()


class HotPath:
    """the user-facing call sequence of train_parsenet_e2e.py:218-277 on our drop-in modules:
    seg-net forward (+ triplet loss) -> NLL -> Evaluation.fitting_loss (mean-shift, match, fit, residual) -> backward"""

    def __init__(self, device, world):
        from src.PointNet import PrimitivesEmbeddingDGCNGn
        from src.residual_utils import Evaluation
        from src.segment_loss import EmbeddingLoss, primitive_loss
        torch.manual_seed(0)
        self.loss = EmbeddingLoss(margin=1.0)
        self.model = PrimitivesEmbeddingDGCNGn(embedding=True, emb_size=EMB, primitives=True, num_primitives=N_PRIM,
                                               loss_function=self.loss.triplet_loss, mode=5, num_channels=6,
                                               nn_nb=KNN_K).to(device)
        self.evaluation = Evaluation(open_decoder=seeded_splinenet(0, 1, device),
                                     closed_decoder=seeded_splinenet(1, 2, device))
        self.primitive_loss = primitive_loss
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=1e-4)
        self.world = world
        self.device = device
        self.clusters = []
        self.codes = cluster_codes(device)

    def step(self, x, lab_np, prim_np, lab, prim):
        """x (B,6,N) cuda; lab_np/prim_np numpy (B,N) for the host-side matching; lab/prim cuda -> loss tensor"""
        self.opt.zero_grad(set_to_none=True)
        emb, lp, el = self.model(x, lab, True)
        loss = el.mean() + self.primitive_loss(lp, prim)
        if FIT_STAGE:
            pts = x[:, 0:3].permute(0, 2, 1).contiguous()
            nrm = x[:, 3:6].permute(0, 2, 1).contiguous()
            emb = pin_clusters(emb, lab, self.codes)
            res, extra = self.evaluation.fitting_loss(emb.permute(0, 2, 1), pts, nrm, lab_np, prim_np.copy(), lp,
                                                      quantile=0.025, iterations=MS_ITERS, lamb=0.1)
            loss = loss + torch.stack([r.reshape(()) for r in res[0::5]]).mean()
            self.clusters.append(len(np.unique(extra[1])))
        loss.backward()
        if self.world > 1:
            from pnb200.parallel import allreduce_mean_grads
            allreduce_mean_grads(self.params, self.world)
        self.opt.step()
        return loss


def run_ours(args):
    from pnb200 import cabi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    hp = HotPath(dev, world)
    B = BATCH_PER_GPU
    from pnb200.input_pipeline import host_rotations
    host = [make_host_batch(B, N_POINTS, seed=100 * rank + i) for i in range(2)]
    # resident batches: preprocessed once, outside the timed region
    dev_batches = []
    for pts_h, nrm_h, lab_h, prim_h in host:
        R = torch.from_numpy(host_rotations(pts_h.numpy())).to(dev)
        dev_batches.append((device_input(pts_h.to(dev), nrm_h.to(dev), R), lab_h.to(dev), prim_h.to(dev)))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    host_np = [(hb[2].numpy(), hb[3].numpy()) for hb in host]

    def resident_step(i):
        x, lab, prim = dev_batches[i % 2]
        np.random.seed(i)
        return hp.step(x, host_np[i % 2][0], host_np[i % 2][1], lab, prim)

    stage_in = [tuple(torch.empty_like(t, device=dev) for t in hb) for hb in host]
    rot_host = [torch.empty((B, 3, 3), dtype=torch.float32).pin_memory() for _ in host]
    rot_dev = [torch.empty((B, 3, 3), dtype=torch.float32, device=dev) for _ in host]
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    losses = []

    def e2e_step(i, last=False):
        """public-API step with HOST inputs: pinned H2D of the batch, the step, D2H of its loss.  The loss of step i is
        copied asynchronously and consumed after step i+1 has been enqueued (lagged logging, like a production loop), so
        the host keeps one step of launch work ahead of the device; the last step's loss is read inside the region."""
        hp_, hn, hl, hpm = host[i % 2]
        pts_d, nrm_d, lab, prim = stage_in[i % 2]            # double-buffered device staging of the inputs
        # input pipeline of the loader: the 3x3 PCA / rotation of every shape on the host (reference arithmetic,
        # ~1 ms per batch, nothing read back), rotation + extent + scaling of the 16 x 10k points on the device
        rot_host[i % 2].numpy()[...] = host_rotations(hp_.numpy())
        pts_d.copy_(hp_, non_blocking=True); nrm_d.copy_(hn, non_blocking=True)
        lab.copy_(hl, non_blocking=True); prim.copy_(hpm, non_blocking=True)
        rot_dev[i % 2].copy_(rot_host[i % 2], non_blocking=True)
        x = device_input(pts_d, nrm_d, rot_dev[i % 2])
        np.random.seed(i)
        loss = hp.step(x, host_np[i % 2][0], host_np[i % 2][1], lab, prim)
        loss_host[i % 2:i % 2 + 1].copy_(loss.detach().reshape(1), non_blocking=True)     # D2H read of the step's result
        loss_ready[i % 2].record()
        if i > 0:
            loss_ready[(i - 1) % 2].synchronize()
            losses.append(float(loss_host[(i - 1) % 2]))
        if last:
            loss_ready[i % 2].synchronize()
            losses.append(float(loss_host[i % 2]))

    # the clock sampler is initialised BEFORE the warm-up (NVML initialisation is the expensive part); its rows are filtered to
    # the timed region afterwards
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("PN_BENCH_NO_SAMPLER"):      # (diagnostic switch; the driver's runs keep the sampler)
        sampler.start()
    for i in range(args.warmup):
        flush.zero_()
        resident_step(i)
        flush.zero_()
        e2e_step(i, last=True)
    # Settling: on fresh boxes the first steps after the warm-up were repeatedly 1.2 - 2.5x slower than the steady state, for
    # 0.1 - 1 s, in the first process of a box only and for 1 as well as for N GPUs (profiles/r02_scaling.md: not the clock
    # sampler, not the first barrier -- something outside the process).  The W warm-up steps are therefore followed by untimed
    # steps until three consecutive step times agree within 5 % (at most 12; the decision is the max over ranks, so every rank
    # runs the same number of steps).  The count is reported as config.settle_steps.
    # Everything that allocates -- the snapshot of the model / optimizer state both timed loops start from, the CUDA events of
    # the timed loops -- is created HERE, before the settle steps: a deep copy taken right before the timed loop left the
    # caching allocator with re-split blocks, and the first timed step then paid for the cudaMallocs (366 / 333 / 255 ms first
    # steps against a 119 ms steady state; the e2e loop, which only restores in place, never showed it).
    def state_tensors():
        ts = list(hp.model.parameters()) + list(hp.model.buffers())
        for st_ in hp.opt.state.values():
            ts += [v for v in st_.values() if torch.is_tensor(v)]
        return ts

    def restore(saved):
        with torch.no_grad():
            for t_, s_ in zip(state_tensors(), saved):
                t_.copy_(s_)

    def new_events(n):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        for e in evs:
            e.record()
        return evs

    snap = [t_.detach().clone() for t_ in state_tensors()]
    cabi.EVENT_POOL[:] = new_events(2 * (MS_ITERS + 2) * args.steps)
    marks_pool = new_events(2 * args.steps + 8)
    barrier()
    settle_ms = []
    while len(settle_ms) < 12:
        a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True)
        flush.zero_()
        a.record()
        resident_step(args.warmup + len(settle_ms))
        b_.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b_)], device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        settle_ms.append(float(t.item()))
        if len(settle_ms) >= 3 and max(settle_ms[-3:]) <= 1.05 * min(settle_ms[-3:]):
            break
    # both timed loops start from the SAME model / optimizer state (the clustering, hence the number and kind of fitted
    # segments, drifts with every Adam step: without this the two loops would time different workloads): restored in place
    restore(snap)
    # ---- timed: resident inputs
    dominant = "pn_ms_iter_fwd_tc" if FIT_STAGE else "pn_knn"
    cabi.TIMED[dominant] = []
    if FIT_STAGE:
        cabi.TIMED["pn_ms_iter_bwd_tc"] = []        # (dense backward: only runs with PN_MS_SPARSE_BWD=0)
        cabi.TIMED["pn_ms_iter_fwd_tma"] = []       # default forward (TMA operands); same roofline entry as the loader-warp kernel
        cabi.TIMED["pn_ms_iter_bwd_tma"] = []
    barrier()
    cabi.reset_launch_count()
    from src.primitive_forward import STATS as fit_stats
    for k in fit_stats:
        fit_stats[k] = 0
    t_wall0 = time.time()
    ev0 = marks_pool.pop(); ev1 = marks_pool.pop()
    ev0.record()
    marks_res = []
    for i in range(args.steps):
        flush.zero_()                           # L2 flush between timed iterations (256 MB > 126 MB L2)
        resident_step(i)
        marks_res.append(marks_pool.pop()); marks_res[-1].record()
        # the host stays at most one step ahead of the device, as in the e2e loop below (where the lagged read of the loss does
        # it): without this bound the un-synchronised loop showed sporadic 1.2 - 1.5x steps on shared 2-GPU boxes
        # (profiles/r02_scaling.md), the e2e loop and the settle steps never did
        if i > 0:
            marks_res[i - 1].synchronize()
    ev1.record()
    barrier()
    ms_res = ev0.elapsed_time(ev1)
    launches = cabi.launch_count()
    fits_per_step = {k: v / args.steps for k, v in fit_stats.items()}
    kern_ms = [a.elapsed_time(b) for a, b in cabi.TIMED.pop(dominant) + cabi.TIMED.pop("pn_ms_iter_fwd_tma", [])]
    bwd_ms = [a.elapsed_time(b) for a, b in cabi.TIMED.pop("pn_ms_iter_bwd_tc", []) + cabi.TIMED.pop("pn_ms_iter_bwd_tma", [])]
    # ---- timed: end to end (H2D of inputs + D2H of the loss inside the region)
    restore(snap)
    barrier()
    ev2 = marks_pool.pop(); ev3 = marks_pool.pop()
    ev2.record()
    mem0 = torch.cuda.memory_stats(dev)
    marks_e2e = []
    for i in range(args.steps):
        flush.zero_()
        e2e_step(i, last=(i == args.steps - 1))
        marks_e2e.append(marks_pool.pop()); marks_e2e[-1].record()
    ev3.record()
    barrier()
    t_wall1 = time.time()
    ms_e2e = ev2.elapsed_time(ev3)
    mem1 = torch.cuda.memory_stats(dev)
    alloc = {"cudaMalloc_calls_in_e2e_region": int(mem1.get("num_device_alloc", 0) - mem0.get("num_device_alloc", 0)),
             "alloc_retries_in_e2e_region": int(mem1.get("num_alloc_retries", 0) - mem0.get("num_alloc_retries", 0)),
             "peak_allocated_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2)}
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    # ---- strong scaling (N > 1 only): the SAME global batch of 16 shapes split over the ranks (BASELINE configs 4 / 5:
    # 4 resp. 2 shapes per GPU), resident inputs, same model state as the other two loops
    strong = None
    if world > 1 and BATCH_PER_GPU % world == 0:
        Bs = BATCH_PER_GPU // world
        restore(snap)

        def strong_step(i):
            x, lab, prim = dev_batches[i % 2]
            np.random.seed(i)
            return hp.step(x[:Bs], host_np[i % 2][0][:Bs], host_np[i % 2][1][:Bs], lab[:Bs], prim[:Bs])

        for i in range(min(args.warmup, 3)):
            strong_step(i)
        barrier()
        ev4 = torch.cuda.Event(enable_timing=True); ev5 = torch.cuda.Event(enable_timing=True)
        ev4.record()
        for i in range(args.steps):
            flush.zero_()
            strong_step(i)
        ev5.record()
        barrier()
        strong = ev4.elapsed_time(ev5)
    per_rank = None
    if world > 1:
        import torch.distributed as dist
        mine = torch.tensor([ms_res, ms_e2e, strong if strong is not None else 0.0], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu().numpy()                 # (world, 3)
        per_rank = {"resident_ms_per_step": [round(float(v) / args.steps, 2) for v in allr[:, 0]],
                    "e2e_ms_per_step": [round(float(v) / args.steps, 2) for v in allr[:, 1]]}
        ms_res, ms_e2e = float(allr[:, 0].max()), float(allr[:, 1].max())
        if strong is not None:
            strong = float(allr[:, 2].max())
    if rank != 0:
        return
    pk = peaks()
    shapes_total = B * world * args.steps
    value = shapes_total / (ms_res / 1e3)
    e2e_v = shapes_total / (ms_e2e / 1e3)
    per_launch_ms = float(np.mean(kern_ms)) if kern_ms else None
    if FIT_STAGE:
        # dominant kernel of the step since the mean-shift backward only visits the centre rows: the fused mean-shift
        # iteration (one launch per iteration, 10 per step).  SURVEY 8(d): 4 N^2 d flop per shape per iteration.
        from pnb200 import meanshift as _pms
        fwd_flop = B * 4.0 * N_POINTS * N_POINTS * EMB
        fwd_ach = fwd_flop / (per_launch_ms * 1e-3) / 1e12 if per_launch_ms else None
        kname = "ms_fwd_tma_kernel<1> (pn_ms_iter_fwd_tma)" if _pms.USE_TMA else "ms_fwd_tc_kernel (pn_ms_iter_fwd_tc)"
        roof = {"kernel": kname + ": one fused mean-shift iteration (S = Y X^T, exp, O += P X, normalise), tcgen05 split-TF32 with "
                          "TMEM accumulators, operand tiles by TMA, batch of %d shapes" % B,
                "bound": "tensor", "achieved": fwd_ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": (fwd_ach / pk["tf_sustained"]) if fwd_ach else None,
                "traffic": MS_FWD_TRAFFIC_BYTES, "traffic_source": MS_FWD_TRAFFIC_SOURCE,
                "algorithmic_flop_per_launch": fwd_flop,
                "peak_source": pk["source"] + " (cuBLAS bf16 dense, sustained)",
                "note": "algorithmic flop = 4*N^2*d per shape per iteration (SURVEY 8d) x 16 shapes per launch; the kernel issues "
                        "3 tf32 MMAs per product (split precision, fp32-accurate result), i.e. the ceiling of this formulation "
                        "is 1/6 of the bf16 peak the fraction is quoted against",
                "launch_ms": per_launch_ms, "launches_timed": len(kern_ms)}
        if bwd_ms:
            roof["dense_backward_launch_ms"] = float(np.mean(bwd_ms))
    else:
        alg_bytes = B * (N_POINTS * 64 * 4 + N_POINTS * KNN_K * 4)
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms else None
        roof = {"kernel": "knn_kernel (pn_knn)", "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": (achieved / pk["hbm_gbs"]) if achieved else None, "traffic": None,
                "peak_source": pk["source"], "launch_ms": per_launch_ms, "launches_timed": len(kern_ms)}
    h2d = int(sum(t.numel() * t.element_size() for t in host[0])) + B * 9 * 4
    out = {
        "metric": "shapes/sec (10k pts, B=16) seg+spline-fit fwd/bwd at 1/2/4/8 B200; Chamfer err",
        "value": value, "unit": "shapes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "train_parsenet_e2e.py step (BASELINE config 5 call sequence at the config-4 batch: 16 x 10k "
                               "pts + normals per GPU, k=80, mode 5): seg-net fwd + triplet/NLL" + (" + Evaluation.fitting_loss "
                               "(mean-shift %d it, match, primitive + open/closed SplineNet fit, residual / Chamfer)" % MS_ITERS
                               if FIT_STAGE else "") + " + bwd + Adam; shapes = 8 patches (plane, sphere, cylinder, cone, "
                               "open spline, closed spline, ...); random-init weights (seg net and SplineNets)",
                   "per_gpu_batch": B, "global_batch": B * world, "n_points": N_POINTS, "knn_k": KNN_K,
                   "parallelism": f"dp{world}", "l2": "256 MB flush write between timed steps; per-step working set "
                                                      ">> 126 MB L2"},
        "e2e": {"value": e2e_v, "unit": "shapes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": (losses[-1] if losses else None),
                "loss_read": "every step, lagged by one step (async D2H + event)",
                "input_pipeline": "raw clouds from pinned host memory; per-shape PCA rotation on the host (3x3), rotation + extent "
                                  "+ scaling of the points / normals on the device (Dataset.get_train, align_canonical=True)",
                "allocator": alloc},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
    }
    out["per_step_ms"] = {"resident": [round(a.elapsed_time(b), 2) for a, b in zip([ev0] + marks_res[:-1], marks_res)],
                          "e2e": [round(a.elapsed_time(b), 2) for a, b in zip([ev2] + marks_e2e[:-1], marks_e2e)]}
    out["config"]["settle_steps"] = len(settle_ms)
    out["config"]["settle_ms"] = [round(v, 1) for v in settle_ms]
    out["config"]["mean_clusters_per_shape_last"] = (float(np.mean(hp.clusters[-4:])) if hp.clusters else None)
    out["config"]["fitted_segments_per_step"] = fits_per_step
    out["config"]["workload_pin"] = ("embedding + %.1f x RMS x code[gt patch]: mean-shift recovers the %d ground-truth patches of "
                                     "every shape at every step (bench.py pin_clusters)" % (PIN_ALPHA, N_PATCHES))
    if per_rank is not None:
        out["per_rank"] = per_rank
    if strong is not None:
        out["strong_scaling"] = {"global_batch": BATCH_PER_GPU, "per_gpu_batch": BATCH_PER_GPU // world,
                                 "ms_per_step": strong / args.steps, "value": BATCH_PER_GPU * args.steps / (strong / 1e3),
                                 "unit": "shapes/s", "note": "same 16 shapes split over the ranks (BASELINE configs 4 / 5), "
                                                             "resident inputs; the headline `value` is weak scaling"}
    if world == 1:
        try:
            out["inference_path"] = inference_path(hp, dev_batches[0], host_np[0], dev)
        except Exception as exc:
            out["inference_path"] = {"value": None, "sample": f"failed: {type(exc).__name__}: {str(exc)[:160]}"}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_subprocess()
        try:
            out["gpu_eager_baseline"] = gpu_eager_baseline(dev)
        except Exception as exc:
            import traceback
            where = " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in traceback.extract_tb(exc.__traceback__)[-4:])
            out["gpu_eager_baseline"] = {"value": None, "unit": "shapes/s",
                                         "sample": f"failed: {type(exc).__name__}: {str(exc)[:160]} at {where}"}
    print(json.dumps(out))


def inference_path(hp, dev_batch, host_np, dev, reps=3):
    """SURVEY 8f-2, second workload (not the headline): the clustering path of generate_predictions.py:131-156 on the same 16
    shapes -- seg-net forward without gradients, normalised embedding, bandwidth, 50 mean-shift iterations, nms (what
    Evaluation.guard_mean_shift does per shape, here through the batched entry points), then the matched-segment IoU metrics on
    the host.  Device time by CUDA events, host metrics included in the region."""
    from pnb200 import meanshift as _ms
    from pnb200.losses import l2_normalize
    from pnb200.assign import match_batched
    from src.segment_utils import SIOU_matched_segments, segment_types_batched
    x, lab, prim = dev_batch
    lab_np, prim_np = host_np
    B = x.shape[0]
    times, ious = [], None
    for rep in range(reps + 1):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        np.random.seed(rep)
        a.record()
        with torch.no_grad():
            emb, lp, _ = hp.model(x, lab, False)
            E = l2_normalize(pin_clusters(emb, lab, hp.codes).permute(0, 2, 1))
            bws = torch.clamp(_ms.compute_bandwidth_batched(E, 10000, 0.015), min=_ms.BW_FLOOR)
            Y = _ms.mean_shift_iters(E, bws, 50)
            members = _ms.nearest_center_batched(E, Y)
            ids, labels_dev, K = _ms.nms_batched(Y, E, bws, members)
            cols = match_batched(labels_dev, lab, 50)[0]          # IoU cost + Hungarian of all shapes on the device
            cluster_np = labels_dev.cpu().numpy()
            onehot = torch.zeros((B, x.shape[2], 64), device=dev).scatter_(2, labels_dev.unsqueeze(2).clamp(max=63), 1.0)
            types = segment_types_batched(torch.max(lp, 1)[1], onehot).cpu().numpy()
            cols = cols.cpu().numpy()
        ious = [SIOU_matched_segments(lab_np[i], cluster_np[i], None, prim_np[i].copy(), None, prim_pred_seg=types[i, :K[i]],
                                      matching=(np.arange(50), cols[i]))[0] for i in range(B)]
        b.record(); torch.cuda.synchronize()
        if rep:
            times.append(a.elapsed_time(b))
    ms = float(np.mean(times))
    return {"value": B / (ms / 1e3), "unit": "shapes/s", "ms_per_batch": ms, "batch": B, "mean_shift_iterations": 50,
            "mean_segment_iou": float(np.mean(ious)),
            "workload": "generate_predictions.py:131-156: seg-net fwd (no grad) + bandwidth + 50 mean-shift iterations + nms + SIOU (IoU cost and "
                        "Hungarian matching of all shapes on the device), "
                        "16 x 10k points"}


FIT_STAGE = True
MS_ITERS = 10
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from the round's `ncu --set full` capture
# (same command, B = 16): filled in from profiles/ when the capture exists, else null
MS_FWD_TRAFFIC_BYTES = 534.8e6
MS_FWD_TRAFFIC_SOURCE = "profiles/r02_ncu_ms_fwd_tma.md (ncu --set full, B = 16, mean of two launches: 459.4 / 610.2 MB)"


# ------------------------------------------------------------------------------------------------ CPU arm
def _port_weights():
    """seeded weights for the oracle port from the committed shape fixture (no import of the product package here)"""
    from oracle.port.common import seeded_state_dict
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")) as f:
        shapes = json.load(f)
    sd = seeded_state_dict(shapes["segnet_mode5_emb128_prim10"], seed=0)
    nets = {"open": seeded_state_dict(shapes["splinenet_open_mode0"], seed=1),
            "closed": seeded_state_dict(shapes["splinenet_closed_mode1"], seed=2)}
    return sd, nets


def _port_one_shape(seed, device=None):
    """ONE shape of the bench workload through the oracle port (torch restatement of the reference path, pinned by
    tests/golden): seg-net forward + triplet/NLL, the reference's Evaluation.fitting_loss (bandwidth + 10 mean-shift iterations
    + nms, Hungarian match, per-segment primitive / SplineNet fits, residuals) on the pinned embedding, backward through
    everything.  device None = CPU tensors (the CPU arm); a cuda device = the same torch ops eagerly on the GPU."""
    import torch.nn.functional as F
    from oracle.port import e2e as pe2e, segnet as port
    from tools.synth import ALL_KINDS, synth_cloud
    sd, nets = _port_weights()
    pts, nrm, lab, prim = synth_cloud(1, N_POINTS, seed=seed, n_patches=N_PATCHES, kinds=ALL_KINDS)
    x = torch.from_numpy(np.concatenate([pts, nrm], 2)).permute(0, 2, 1).contiguous()
    tp, tn, tprim, tlab = torch.from_numpy(pts[0]), torch.from_numpy(nrm[0]), torch.from_numpy(prim), torch.from_numpy(lab)
    codes = cluster_codes()
    if device is not None:
        sd = {k: v.to(device) for k, v in sd.items()}
        nets = {n: {k: v.to(device) for k, v in d.items()} for n, d in nets.items()}
        x, tp, tn, tprim, tlab, codes = (t.to(device) for t in (x, tp, tn, tprim, tlab, codes))
    sd = {n: v.detach().clone().requires_grad_(v.is_floating_point()) for n, v in sd.items()}
    sync = (lambda: torch.cuda.synchronize(device)) if device is not None else (lambda: None)
    import contextlib
    # (inputs and seeded weights above are built on the CPU: their generators are CPU generators; only the port's own
    # factory calls -- torch.zeros / eye / arange ... -- follow the device of the run)
    with (torch.device(device) if device is not None else contextlib.nullcontext()):
        sync()
        t0 = time.time()
        emb, lp, _, _, _ = port.segnet_fwd(sd, x, KNN_K, 5)
        np.random.seed(0)
        el = port.triplet_loss(emb, lab, 1.0)
        nll = F.nll_loss(lp, tprim)
        sync()
        t_seg = time.time() - t0
        loss = el.sum() + nll
        emb = pin_clusters(emb, tlab, codes)
        fl, _, dist, cl = pe2e.fitting_loss(emb[0].t(), tp, tn, lab[0], prim[0].copy(), nets, 0.025, MS_ITERS, 0.1)
        loss = loss + fl[0].reshape(())
        loss.backward()
        sync()
        dt = time.time() - t0
    kinds = sorted(v[0] for v in dist.values())
    return dt, (f"1 shape x {N_POINTS} pts, k={KNN_K}: seg-net fwd+losses ({t_seg:.1f} s), Evaluation.fitting_loss (mean-shift "
                f"{MS_ITERS} it + nms, match, {len(dist)} segment fits {kinds}, residuals; {len(np.unique(cl))} clusters), bwd; "
                f"{dt:.1f} s wall")


def cpu_baseline():
    """the reference's CPU path = the oracle port on this host's cores (the reference itself is Python under /root/reference
    and cannot travel to the GPU box; the port reproduces its numbers to 1e-4 .. 1e-6 on the golden fixtures)"""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dt, sample = _port_one_shape(seed=0)
    return {"value": 1.0 / dt, "unit": "shapes/s", "cores": cores, "threads": torch.get_num_threads(), "kind": "port",
            "sample": sample}


def cpu_baseline_subprocess(timeout_s=420):
    """run the CPU arm in a fresh interpreter: no CUDA context, no product library, and -- what made the N = 1 line of the
    round-1 scaling run take 453 s -- no inherited OMP_NUM_THREADS=1 from a torchrun launcher pinning the port to one core"""
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "WORLD_SIZE",
                                                             "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                           env=env, capture_output=True, text=True, timeout=timeout_s)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as exc:              # a failure of the CPU arm must not cost the measured GPU line
        return {"value": None, "unit": "shapes/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"failed: {type(exc).__name__}: {str(exc)[:200]}"}


def gpu_eager_baseline(dev):
    """SURVEY 8(d) "the real bar": the same torch ops as the CPU arm, run eagerly on this B200 (the port with its tensors on
    the device; kNN as the reference's matmul + topk).  torch factory calls inside the port follow the default device; the few
    numpy round trips are patched to hop through the host for the duration of the call."""
    from oracle.port import segnet as port

    def knn_torch(x, k, metric):
        with torch.no_grad():
            outs = []
            for b in range(x.shape[0]):
                if metric == 0:
                    xb = x[b:b + 1]
                    inner = -2 * torch.matmul(xb.transpose(2, 1), xb)
                    xx = torch.sum(xb ** 2, dim=1, keepdim=True)
                    d = -xx - inner - xx.transpose(2, 1)
                else:
                    p, n = x[b:b + 1, 0:3], x[b:b + 1, 3:6]
                    xx = torch.sum(p ** 2, dim=1, keepdim=True)
                    pd = xx - 2 * torch.matmul(p.transpose(2, 1), p) + xx.transpose(2, 1)
                    d = -(pd * (1 + (2 - 2 * torch.matmul(n.transpose(2, 1), n))))
                outs.append(d.topk(k=k, dim=-1)[1])
            return torch.cat(outs, 0)

    saved = (port.knn_idx, torch.from_numpy, torch.Tensor.numpy, torch.as_tensor)
    # factory calls without a device argument follow the run's device -- patched on the torch module itself, because the
    # backward of the port's autograd Functions runs on the autograd engine's thread, where a torch.device(...) context
    # (thread-local) is not active
    factories = {n: getattr(torch, n) for n in ("zeros", "ones", "eye", "full", "arange", "tensor", "empty", "linspace")}

    def on_dev(fn):
        def wrapped(*a, **k):
            if k.get("device") is None:
                k["device"] = dev
            return fn(*a, **k)
        return wrapped

    try:
        for n, fn in factories.items():
            setattr(torch, n, on_dev(fn))
        port.knn_idx = knn_torch
        torch.from_numpy = lambda a: saved[1](a).to(dev)
        torch.Tensor.numpy = lambda self, *a, **k: saved[2](self.detach().cpu(), *a, **k)
        torch.as_tensor = lambda a, *ar, **k: (saved[3](a, *ar, **k).to(dev) if isinstance(a, np.ndarray)
                                               else saved[3](a, *ar, **k))
        _port_one_shape(seed=1, device=dev)                      # warm-up (cuBLAS / cuSOLVER handles, allocator)
        dt, sample = _port_one_shape(seed=0, device=dev)
    finally:
        port.knn_idx, torch.from_numpy, torch.Tensor.numpy, torch.as_tensor = saved
        for n, fn in factories.items():
            setattr(torch, n, fn)
    return {"value": 1.0 / dt, "unit": "shapes/s", "kind": "port on cuda (torch eager, reference formulation)", "sample": sample}


def run_reference(args):
    """reference arm: the reference's CPU implementation of the path = oracle port on the host cores
    (the reference itself is Python under /root/reference and cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if os.environ.get("OMP_NUM_THREADS") == "1" and (os.cpu_count() or 1) > 1:
        # torchrun pins OMP_NUM_THREADS=1 for every worker; the CPU arm is ONE process that should use the whole host
        return print(json.dumps(cpu_reference_line(args, cpu_baseline_subprocess())))
    vals = []
    base = None
    for _ in range(max(1, min(args.steps, 2))):
        base = cpu_baseline()
        vals.append(base["value"])
    base["value"] = float(np.mean(vals))
    print(json.dumps(cpu_reference_line(args, base)))


def cpu_reference_line(args, base):
    v = base["value"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return {
        "impl": "reference",
        "metric": "shapes/sec (10k pts, B=16) seg+spline-fit fwd/bwd at 1/2/4/8 B200; Chamfer err",
        "value": v, "unit": "shapes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": (1e3 / v * BATCH_PER_GPU) if v else None, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "same as the ours arm (same synthetic generator, same embedding pin -> the same 8 fitted segments "
                               "per shape); each step is a bounded sample (1 shape) of the 16-shape batch, scaled linearly",
                   "per_gpu_batch": BATCH_PER_GPU, "n_points": N_POINTS, "knn_k": KNN_K},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
