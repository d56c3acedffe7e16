#!/usr/bin/env python
"""Benchmark of the FCOS detection hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "pass" = the inference post-process (K1 score -> K2 top-k -> K3 NMS + clip) over one batch of synthetic head
outputs at BASELINE config 2 (COCO 832x1344, 80 classes, batch 16 per GPU).  A "step" = PASSES_PER_STEP (128)
consecutive passes (16 replays of a CUDA graph of 8 passes over 8 rotating input sets), so that the timed region is
>= 50 ms whatever --steps is and device-side launch skew between ranks is negligible against it.  ONE JSON line:
  value         whole-job img/s with inputs resident in HBM (8 input sets of 127 MB rotated: every pass misses L2)
  e2e           same metric through the public API (FCOSHead.detect) from pinned HOST buffers, H2D of the head
                outputs and D2H of the detections inside the timed region
  roofline      K1 (score_points, the kernel that moves > 95 % of the bytes) timed alone with CUDA events
  cpu_baseline  the oracle port of the reference's CPU path (torch CPU ops + torchvision.ops.batched_nms, exactly
                the call of head.py:94) on this host, full batch of 16
  config3       BASELINE config 3: target assignment + GIoU fwd/bwd, B=32, M<=100: us/batch, roofline, cpu_baseline
  config4       dense crowd: 5 000-candidate NMS and 300-GT assignment (time only: latency-bound)
  config5       B=256 STRONG-scaled over the ranks: post-process + FCOSTargetLoss step + one gather of the detections
                and the loss sums
  reference_eager_b200   the reference's torch ops + torchvision CUDA NMS run eagerly on this GPU (informative)
With --impl reference only the CPU arm runs (rank 0), without importing the product package.
Multi-GPU (torchrun): batch sharded by rank (weak scaling, 16 images per GPU per pass); the packed detections of
every graph of 8 passes are gathered to rank 0 through NVLink peer memory (sharding.PeerGather; NCCL all_gather where
symmetric memory is unavailable), overlapped with the following graphs; time = max over ranks.
"""
import argparse
import importlib.util
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# 8 batches in flight + 2 replay streams + the gather stream exceed the default 8 hardware work queues; streams that
# share a queue pick up false dependencies (measured at 2 GPUs: 24.2 -> 22.5 us per pass with 32 queues)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402


def _load_workloads():
    """The synthetic-input definitions, loaded BY PATH: importing the package would map libb200det.so into the
    process, and the reference arm must not touch the product."""
    path = os.path.join(ROOT, "pytorch_object_detection_b200", "workloads.py")
    spec = importlib.util.spec_from_file_location("b200det_bench_workloads", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


W = _load_workloads()

BATCH = 16
NCLS = 80
MAX_BOX = 1000
SCORE_THR = 0.05
NMS_THR = 0.6
SETS = 8                    # input sets rotated through (each 127 MB; L2 is 126 MB)
ROUNDS_PER_STEP = 16        # graph replays per step
PASSES_PER_STEP = SETS * ROUNDS_PER_STEP
TRAIN_BATCH = 32
TRAIN_MAX_GT = 100
CROWD_N, CROWD_B, CROWD_GT = 5000, 8, 300
STRONG_BATCH = 256
P = W.num_points(W.COCO_LEVELS)
WORKLOAD = f"COCO 832x1344 FCOS post-process (score+top-k {MAX_BOX}+NMS {NMS_THR}+clip), {NCLS} classes, " \
           f"P={P}, batch {BATCH} per GPU"
METRIC = "postprocess_throughput"


def bench_config(world):
    """The `config` object: identical for the b200 and the reference arm at the same N."""
    return {"workload": WORKLOAD, "batch_per_pass_per_gpu": BATCH, "global_batch": world * BATCH,
            "passes_per_step": PASSES_PER_STEP,
            "l2": f"{SETS} input sets of 127 MB rotated, every pass reads a set that left the 126 MB L2"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f)[kernel]
        return float(rec["dram_read_bytes"]) + float(rec["dram_write_bytes"]), f"profiles/ncu_traffic.json ({rec['from']})"
    except Exception:                                       # noqa: BLE001
        return None, "no ncu capture committed for this kernel"


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:                                   # noqa: BLE001
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:                       # noqa: BLE001
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:                           # noqa: BLE001
                    pass
            time.sleep(0.002)

    def __enter__(self):
        self._active.set()
        return self

    def __exit__(self, *a):
        self._active.clear()

    def summary(self):
        self._stop.set()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_postprocess_leg(steps, warmup):
    """FCOSHead + ClipBoxes of the reference on the CPU: the oracle port (torch CPU ops in the reference's order)
    with torchvision.ops.batched_nms — the very call of head.py:94 — on the FULL batch of 16, per step."""
    import torchvision
    from oracle import fcos_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    nms = torchvision.ops.batched_nms
    xs = [W.head_outputs(BATCH, NCLS, W.COCO_LEVELS, seed=1000 + s) for s in range(2)]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            x = xs[i % 2]
            t0 = time.perf_counter()
            dets = O.detect(x, SCORE_THR, NMS_THR, MAX_BOX, W.STRIDES, nms_fn=nms)
            for d in dets:
                O.clip_boxes_(d[2], *W.COCO_HW)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    thr = torch.get_num_threads()
    return {"value": BATCH * len(times) / total, "unit": "img/s", "cores": thr, "kind": "port",
            "ms_per_step": 1e3 * total / len(times), "best_img_s": BATCH / min(times),
            "sample": f"full batch of {BATCH} images per step x {len(times)} steps after {warmup} warm-up, oracle port "
                      f"(torch CPU ops + torchvision.ops.batched_nms as head.py:94), {thr} torch threads, "
                      f"{os.cpu_count()} host cores"}


def cpu_train_leg(sample_images=8, reps=2):
    """FCOSGenTargets + compute_reg_loss('giou') forward + backward of the reference on the CPU (oracle port) on a
    bounded sample of config 3; scaled to the batch of 32 (the work is per image)."""
    from oracle import fcos_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    gt, labels = W.gt_boxes(TRAIN_BATCH, TRAIN_MAX_GT, W.COCO_HW, NCLS, seed=3000)
    gt, labels = gt[:sample_images], labels[:sample_images]
    g = torch.Generator().manual_seed(3001)
    regs = [torch.exp(torch.randn(sample_images, 4, h, w, generator=g) + 3).requires_grad_(True) for h, w in W.COCO_LEVELS]
    times = []
    for i in range(reps + 1):
        for t in regs:
            t.grad = None
        t0 = time.perf_counter()
        tgt = O.assign_targets(W.COCO_LEVELS, gt, labels, W.STRIDES, W.HISFCOS_RANGES)
        t1 = time.perf_counter()
        mask = (tgt[1] > -1).squeeze(-1)
        O.reg_loss(regs, tgt[2], mask, "giou").mean().backward()
        t2 = time.perf_counter()
        if i >= 1:
            times.append((t2 - t0, t1 - t0))
    tot = min(t[0] for t in times)
    asg = min(t[1] for t in times)
    scale = TRAIN_BATCH / sample_images
    thr = torch.get_num_threads()
    return {"value": 1e6 * tot * scale, "unit": "us/batch", "cores": thr, "kind": "port",
            "assign_us_per_batch": 1e6 * asg * scale,
            "sample": f"{sample_images} of the {TRAIN_BATCH} images (x {scale:g}), best of {len(times)} after 1 warm-up, "
                      f"oracle port of head.py:235-316 + loss.py:116-177 fwd+bwd, {thr} torch threads"}


def main_reference(args, rank, result_out):
    if rank != 0:
        return
    warmup = max(args.warmup, 3)
    leg = cpu_postprocess_leg(args.steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": leg["value"], "unit": "img/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup, "ms_per_step": leg["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": bench_config(args.gpus),
            "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "step_note": f"a reference step is a bounded sample: ONE pass over a batch of {BATCH} (the b200 arm's step "
                         f"is {PASSES_PER_STEP} passes); img/s is comparable"}
    result_out.write(json.dumps(line) + "\n")
    result_out.flush()


# ------------------------------------------------------------------------------------------------------------
# the b200 arm
# ------------------------------------------------------------------------------------------------------------
def median(v):
    s = sorted(v)
    return s[len(s) // 2]


def main_b200(args, rank, world, local, result_out):
    assert torch.cuda.is_available(), "bench.py needs a GPU: b200det has no CPU path"
    import pytorch_object_detection_b200 as B
    from pytorch_object_detection_b200 import ops, sharding
    warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    peak, peak_src = peaks()
    sampler = ClockSampler(local)

    # ---- synthetic inputs: SETS independent batches, rotated so every pass misses L2 -------------------
    # (two sets are made on the host — the e2e leg copies them from pinned memory every step —, the others with
    # the same distributions directly on the device: 1 GB of host random numbers per rank is only start-up time)
    host_sets = [W.head_outputs(BATCH, NCLS, W.COCO_LEVELS, seed=2000 + 17 * rank + s) for s in range(2)]
    dev_sets = [[[t.to(dev) for t in part] for part in hs] for hs in host_sets]
    for s_i in range(len(host_sets), SETS):
        g_dev = torch.Generator(device=dev).manual_seed(2000 + 17 * rank + s_i)
        cls_d, cnt_d, reg_d = [], [], []
        for h, w in W.COCO_LEVELS:
            cls_d.append(torch.randn(BATCH, NCLS, h, w, device=dev, generator=g_dev) - 4.595)
            cnt_d.append(torch.randn(BATCH, 1, h, w, device=dev, generator=g_dev))
            reg_d.append(torch.exp(torch.randn(BATCH, 4, h, w, device=dev, generator=g_dev) + 3.0))
        dev_sets.append([cls_d, cnt_d, reg_d])
    in_bytes = sum(t.numel() * 4 for part in host_sets[0] for t in part)
    head = B.FCOSHead(SCORE_THR, NMS_THR, MAX_BOX, W.STRIDES)

    # A "round" = one pass over each of the SETS input batches, captured as ONE CUDA graph so the timed loop is not
    # bound by Python launch overhead.  Batches are independent, so inside the graph they are forked round-robin onto
    # `streams` capture streams: K1 (HBM-bound, all SMs) of one batch overlaps the per-image select/NMS kernel (one
    # CTA per image, latency-bound) of the others.  Multi-GPU: the path's only collective is the final gather of the
    # detections.  The packed outputs of a whole round sit in one buffer, so ONE gather serves SETS passes; four
    # output buffers rotate over two streams so that a gather overlaps the kernels of the following rounds.
    k_out = min(MAX_BOX, P)
    pk_bytes = ops.packed_nbytes(BATCH, k_out)
    for hs in dev_sets:                                   # warm the allocator / library before capture
        head.detect(hs, clip_hw=W.COCO_HW)
    torch.cuda.synchronize()

    # The gather: peer-memory pushes on the copy engines + a device-side barrier (sharding.PeerGather); NCCL
    # all_gather when symmetric memory is not available (B200DET_BENCH_GATHER=nccl forces it).  The detections
    # are gathered to rank 0, where an evaluation collects them (Test_coco.py:144-168 writes one result file);
    # B200DET_BENCH_GATHER=peer_all gives every rank every detection instead.
    peer = None
    gather_check = {"ok": None}                             # peer gather compared with NCCL all_gather once
    gather_kind = "none"
    gather_root = None
    if dist is not None:
        gather_kind = "NCCL all_gather"
        gather_mode = os.environ.get("B200DET_BENCH_GATHER", "peer_root")      # peer_root | peer_all | nccl
        gather_root = 0 if gather_mode == "peer_root" else None
        if gather_mode.startswith("peer"):
            try:
                peer = sharding.PeerGather(SETS * pk_bytes, 4, dev)
                gather_kind = ("gather to rank 0" if gather_root == 0 else "all-gather") + \
                    " by peer-memory pushes (copy engines) + device barrier"
            except Exception as e:                          # noqa: BLE001
                print(f"[bench] rank {rank}: symmetric memory unavailable ({type(e).__name__}: {e}); NCCL all_gather",
                      file=sys.stderr)
                peer = None
        flags = torch.tensor([1.0 if peer is not None else 0.0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)        # every rank must take the same path
        if float(flags[0]) == 0.0:
            peer, gather_kind = None, "NCCL all_gather"

    def capture_round(n_streams):
        out_big = torch.empty((SETS, pk_bytes), dtype=torch.uint8, device=dev)
        side = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cs = torch.cuda.current_stream()
            for st in side:
                st.wait_stream(cs)                        # fork
            for s_i, hs in enumerate(dev_sets):
                with torch.cuda.stream(side[s_i % n_streams]):
                    head.detect(hs, clip_hw=W.COCO_HW, out_packed=out_big[s_i])
            for st in side:
                cs.wait_stream(st)                        # join
        # every buffer the timed loop touches exists before it starts (nothing is allocated inside e0..e1)
        full = torch.empty((world, SETS * pk_bytes), dtype=torch.uint8, device=dev) if world > 1 and peer is None else None
        return g, out_big, full

    def timed_rounds(n_streams, steps, warm_steps):
        """`steps` steps of ROUNDS_PER_STEP graph replays each, after `warm_steps` untimed steps through the SAME
        graphs, gather path and streams.  Returns ms per step (max over ranks)."""
        n_buf = 4                                                  # output buffers (graphs) in rotation
        rounds = [capture_round(n_streams) for _ in range(n_buf)]
        pending = [None] * n_buf
        done = [None] * n_buf
        # rounds are replayed alternately on two streams so that the tail of one overlaps the head of the next
        outer = [torch.cuda.Stream(device=dev) for _ in range(2 if n_streams > 1 else 1)]
        gather_stream = torch.cuda.Stream(device=dev) if dist is not None else None
        no_gather = bool(os.environ.get("B200DET_BENCH_NO_GATHER"))                # (debug switch)

        def run_round(r):
            q = r % n_buf
            g, out_big, full = rounds[q]
            with torch.cuda.stream(outer[q % len(outer)]):
                if pending[q] is not None:
                    torch.cuda.current_stream().wait_event(pending[q])   # last gather of this buffer finished
                g.replay()
                done[q] = torch.cuda.Event()
                done[q].record()
            if dist is not None and not no_gather:
                with torch.cuda.stream(gather_stream):      # the collective has its own stream; it waits for this round only
                    gather_stream.wait_event(done[q])
                    if peer is not None:
                        peer.gather(q % peer.slots, out_big.reshape(-1), root=gather_root)
                    else:
                        dist.all_gather_into_tensor(full, out_big.reshape(-1))
                    pending[q] = torch.cuda.Event()
                    pending[q].record()

        def drain():
            cur = torch.cuda.current_stream()
            for q in range(n_buf):
                if done[q] is not None:
                    cur.wait_event(done[q])
                if pending[q] is not None:
                    cur.wait_event(pending[q])
                    pending[q] = None

        for r in range(max(n_buf, warm_steps * ROUNDS_PER_STEP)):
            run_round(r)
        drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler:
            e0.record()
            for st in outer:
                st.wait_event(e0)                         # nothing starts before the start event
            for r in range(steps * ROUNDS_PER_STEP):
                run_round(r)
            drain()                                       # the last gathers are inside the timed region
            e1.record()
            barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        if peer is not None and gather_check["ok"] is None:
            # outside the timed region, once: what the peer-memory gather delivered to its receivers must be what
            # an NCCL all_gather of the same buffers delivers (every rank takes part in both)
            try:
                g, out_big, _ = rounds[0]
                g.replay()
                ref = torch.empty((world, out_big.numel()), dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(ref, out_big.reshape(-1))
                got = peer.gather(0, out_big.reshape(-1), root=gather_root)
                torch.cuda.synchronize()
                same = torch.tensor([1.0], device=dev)
                if gather_root is None or rank == gather_root:
                    same[0] = 1.0 if torch.equal(got, ref) else 0.0
                dist.all_reduce(same, op=dist.ReduceOp.MIN)
                gather_check["ok"] = bool(float(same[0]) == 1.0)
            except Exception as e:                          # noqa: BLE001  (never lose the measurement to the check)
                print(f"[bench] rank {rank}: gather check failed to run: {type(e).__name__}: {e}", file=sys.stderr)
                gather_check["ok"] = False
        return ms / steps

    # ---- value: device-resident, `n_streams` batches in flight; and one batch at a time ------------
    n_streams = max(1, min(args.streams, SETS))
    ms_per_step = timed_rounds(n_streams, args.steps, warmup)
    value = world * BATCH * PASSES_PER_STEP / (ms_per_step * 1e-3)
    ms_few = timed_rounds(min(3, n_streams), max(3, args.steps // 4), 3)
    ms_single = timed_rounds(1, max(3, args.steps // 4), 3)
    launches = args.steps * PASSES_PER_STEP * 2           # score_points + fused select/NMS kernel per pass

    # ---- e2e: pinned host inputs -> H2D -> public API -> D2H of the detections -----------------
    # Two staging slots: the H2D copies of step i+1 (split over two copy streams) overlap the kernels and
    # the D2H of step i; the host waits for every step's packed result in pinned memory.
    def e2e_leg(dtype):
        pinned = [[[t.to(dtype).pin_memory() for t in part] for part in hs] for hs in host_sets[:2]]
        h2d = sum(t.numel() * t.element_size() for part in pinned[0] for t in part)
        e2e_steps = max(3, args.steps)
        slots = []
        for _ in range(2):
            stage = [[torch.empty_like(t, device=dev) for t in part] for part in pinned[0]]
            out_k = torch.empty((pk_bytes,), dtype=torch.uint8, device=dev)
            head.detect(stage, clip_hw=W.COCO_HW, out_packed=out_k)
            torch.cuda.synchronize()
            g_k = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_k):
                head.detect(stage, clip_hw=W.COCO_HW, out_packed=out_k)
            slots.append({"stage": stage, "out": out_k, "graph": g_k,
                          "host": torch.empty((pk_bytes,), dtype=torch.uint8).pin_memory(),
                          "full": torch.empty((world, pk_bytes), dtype=torch.uint8, device=dev) if world > 1 else None,
                          "done": None})
        copy_streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        main = torch.cuda.current_stream()

        def issue(i):
            sl = slots[i % 2]
            src = pinned[i % len(pinned)]
            pairs = [(a, b) for ps, pd in zip(src, sl["stage"]) for a, b in zip(ps, pd)]
            for j, cst in enumerate(copy_streams):
                if sl["done"] is not None:
                    cst.wait_event(sl["done"])                 # the slot's previous step has been consumed
                with torch.cuda.stream(cst):
                    for a, b in pairs[j::2]:
                        b.copy_(a, non_blocking=True)          # H2D of this step's head outputs
                main.wait_stream(cst)
            sl["graph"].replay()
            if dist is not None:
                dist.all_gather_into_tensor(sl["full"], sl["out"])
            sl["host"].copy_(sl["out"], non_blocking=True)     # D2H of the step's detections
            ev = torch.cuda.Event()
            ev.record(main)
            sl["done"] = ev

        def run(n):
            for i in range(n):
                issue(i)
                if i >= 1:
                    slots[(i - 1) % 2]["done"].synchronize()   # host has step i-1's result
            slots[(n - 1) % 2]["done"].synchronize()

        run(3)
        barrier()
        with sampler:
            t0 = time.perf_counter()
            run(e2e_steps)
            barrier()
            e2e_s = max_over_ranks(time.perf_counter() - t0)
        return {"value": world * BATCH * e2e_steps / e2e_s, "unit": "img/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": pk_bytes, "steps": e2e_steps,
                "step": f"one pass over a batch of {BATCH} per GPU",
                "api": "FCOSHead.detect(clip_hw=...) on pinned host inputs"}

    e2e = e2e_leg(torch.float32)
    e2e_half = None
    if os.environ.get("B200DET_BENCH_HALF_E2E", "1") == "1" and getattr(ops, "NATIVE_HALF_POSTPROCESS", False):
        e2e_half = e2e_leg(torch.float16)
        e2e_half["note"] = "fp16 head outputs (autocast, train.py:175) read natively by K1/K2: half the PCIe bytes"

    # ---- config 5: B=256 strong-scaled over the ranks (post-process + training step + collectives) -------
    config5 = strong_scaling_leg(B, ops, sharding, dist, dev, rank, world, barrier, max_over_ranks, sampler)

    # ---- roofline of the dominant kernel (K1) timed alone, same inputs, same rotation ----------
    k1_bytes = BATCH * P * ((NCLS + 1) * 4 + 4 + 2)       # cls + cnt planes read, score f32 + class i16 written
    for hs in dev_sets:
        ops.score_points(hs[0], hs[1], W.STRIDES)
    torch.cuda.synchronize()
    chunk = 2 * SETS                                        # launches per graph replay / event pair
    k1_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(k1_graph):
        for i in range(chunk):
            hs = dev_sets[i % SETS]
            ops.score_points(hs[0], hs[1], W.STRIDES)
    for i in range(3):
        k1_graph.replay()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    with sampler:
        for a, b in evs:
            a.record()
            k1_graph.replay()
            b.record()
        torch.cuda.synchronize()
    k1_ms = sorted(a.elapsed_time(b) / chunk for a, b in evs)
    k1_avg = sum(k1_ms) / len(k1_ms)
    achieved = k1_bytes / (k1_avg * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic("score_points_kernel")
    roofline = {"kernel": "score_points_kernel (K1)", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "bytes_per_launch": k1_bytes, "us_per_launch": 1e3 * k1_avg,
                "us_median": 1e3 * median(k1_ms), "event_pairs": len(evs), "launches_per_pair": chunk}

    config3 = config4 = eager = cpu = None
    if world == 1:
        config3, config4 = train_legs(B, ops, dev, peak, sampler)
        eager = eager_reference_leg(dev)
        cpu = cpu_postprocess_leg(steps=5, warmup=2)
        cpu3 = cpu_train_leg()
        config3["cpu_baseline"] = {k: cpu3[k] for k in ("value", "unit", "cores", "kind", "sample", "assign_us_per_batch")}
        if config5 is not None:
            config5["cpu_baseline_estimate_ms"] = 1e3 * STRONG_BATCH / cpu["value"] + \
                1e-3 * cpu3["value"] * STRONG_BATCH / TRAIN_BATCH
            config5["cpu_baseline_note"] = "256 / (config-2 CPU img/s) + 8 x (config-3 CPU time per batch of 32)"

    clocks = sampler.summary()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cfg = bench_config(world)
    line = {"metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "us_per_pass": 1e3 * ms_per_step / PASSES_PER_STEP,
            "timing": {"in_flight": f"{n_streams} batches on {n_streams} CUDA streams inside one CUDA graph of {SETS} passes; "
                                    f"a step = {ROUNDS_PER_STEP} replays",
                       "collective": f"one gather of the packed detections per {SETS} passes, overlapped: {gather_kind}"
                                     if world > 1 else "none",
                       "gather_equals_nccl_all_gather": gather_check["ok"],
                       "timed_region_ms": ms_per_step * args.steps},
            "few_in_flight": {"streams": min(3, n_streams), "value": world * BATCH * PASSES_PER_STEP / (ms_few * 1e-3),
                              "unit": "img/s", "us_per_pass": 1e3 * ms_few / PASSES_PER_STEP},
            "single_stream": {"value": world * BATCH * PASSES_PER_STEP / (ms_single * 1e-3), "unit": "img/s",
                              "us_per_pass": 1e3 * ms_single / PASSES_PER_STEP},
            "e2e": e2e, "e2e_fp16_inputs": e2e_half,
            "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None,
            "config3": config3, "config4": config4, "config5": config5, "reference_eager_b200": eager,
            "clocks": clocks}
    result_out.write(json.dumps(line) + "\n")
    result_out.flush()
    if dist is not None:
        dist.destroy_process_group()


def timed_graph(fn, reps, pairs=5, group=4):
    """us per call of fn(i): `group` calls captured as one CUDA graph (no Python between launches), `pairs` event
    pairs of reps/group replays each; returns (median, min) over the pairs."""
    for i in range(group):
        fn(i)
    torch.cuda.synchronize()
    n = max(1, reps // group)
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(group):
                fn(i)
        run = g.replay
    except Exception:                                       # noqa: BLE001  (capture refused: time eagerly)
        torch.cuda.synchronize()
        run = lambda: [fn(i) for i in range(group)]         # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(pairs)]
    for a, b in evs:
        a.record()
        for _ in range(n):
            run()
        b.record()
    torch.cuda.synchronize()
    us = sorted(1e3 * a.elapsed_time(b) / (group * n) for a, b in evs)
    return median(us), us[0]


def train_legs(B, ops, dev, peak, sampler):
    """BASELINE config 3 (target assignment + GIoU fwd/bwd, B=32, M<=100) and config 4 (dense crowd)."""
    reps = 200
    gt, labels = W.gt_boxes(TRAIN_BATCH, TRAIN_MAX_GT, W.COCO_HW, NCLS, seed=3000)
    gt, labels = gt.to(dev), labels.to(dev)
    regs = [[torch.exp(torch.randn(TRAIN_BATCH, 4, h, w, device=dev) + 3).requires_grad_(True)
             for h, w in W.COCO_LEVELS] for _ in range(4)]
    cnts = [[torch.randn(TRAIN_BATCH, 1, h, w, device=dev).requires_grad_(True) for h, w in W.COCO_LEVELS]
            for _ in range(4)]
    gen = B.FCOSGenTargets(W.STRIDES, W.HISFCOS_RANGES)
    fake = [torch.empty(TRAIN_BATCH, 1, h, w, device="meta") for h, w in W.COCO_LEVELS]
    fused = B.FCOSTargetLoss(W.STRIDES, W.HISFCOS_RANGES, "giou")
    clss = []            # two sets of class logits (238 MB each), made when the full step is timed

    def train_step(i):
        """Fused path: targets + GIoU loss + its gradients behind autograd (FCOSTargetLoss)."""
        for t in regs[i % 4]:
            t.grad = None                              # optimizer.zero_grad(set_to_none=True)
        loss, _ = fused.box_cnt_losses(None, regs[i % 4], gt, labels)
        loss.backward()

    def train_step_cnt(i):
        """Same with the centerness BCE branch (loss.py:29-57) in the same launches."""
        for t in regs[i % 4] + cnts[i % 4]:
            t.grad = None
        reg_loss, cnt_loss = fused.box_cnt_losses(cnts[i % 4], regs[i % 4], gt, labels)
        (reg_loss + cnt_loss).backward()

    def train_step_unfused(i):
        """The drop-in modules one by one: FCOSGenTargets -> compute_reg_loss -> backward."""
        for t in regs[i % 4]:
            t.grad = None
        tgt = gen([[fake, fake, fake], gt, labels])
        B.compute_reg_loss(regs[i % 4], tgt[2], None, "giou", _mask_src=tgt[1]).mean().backward()

    def train_step_full(i):
        """The whole FCOSGenTargets + FCOSLoss step: fused assign/box/centerness launch + one focal launch that
        writes the loss and the 238 MB class gradient from one read of the logits."""
        for t in regs[i % 4] + cnts[i % 4] + clss[i % 2]:
            t.grad = None
        fused([(clss[i % 2], cnts[i % 4], regs[i % 4]), gt, labels])[3].backward()

    with sampler:
        us, us_min = timed_graph(train_step, reps)
        us_cnt, _ = timed_graph(train_step_cnt, reps)
        us_unfused, _ = timed_graph(train_step_unfused, reps)
        us_kernels, us_kernels_min = timed_graph(
            lambda i: ops.assign_loss_fused(regs[i % 4], None, W.STRIDES, W.HISFCOS_RANGES, gt, labels, 1), reps)
        us_assign, us_assign_min = timed_graph(
            lambda i: ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt, labels), reps)
        clss.extend([(torch.randn(TRAIN_BATCH, NCLS, h, w, device=dev) - 4.595).requires_grad_(True)
                     for h, w in W.COCO_LEVELS] for _ in range(2))
        us_full, _ = timed_graph(train_step_full, reps)
        us_focal, _ = timed_graph(lambda i: ops.cls_loss_step(clss[i % 2], fused.targets[0],
                                                              num_pos=fused.per_image["num_pos"]), reps)
        cls_h = [[t.detach().half() for t in clss[i]] for i in range(2)]
        scale_state = torch.tensor([65536.0, 0.0], device=dev)
        us_focal_h, _ = timed_graph(lambda i: ops.cls_loss_step(cls_h[i % 2], fused.targets[0],
                                                                num_pos=fused.per_image["num_pos"],
                                                                up_mean=scale_state), reps)
        del cls_h
    del clss[:]
    assign_bytes = TRAIN_BATCH * P * 28 + TRAIN_BATCH * TRAIN_MAX_GT * 24
    # targets 28 B/pt written + reg predictions 16 B/pt in + reg gradients 16 B/pt out = 44.7 MB; the
    # kernel reads predictions at positives only, so its own traffic is targets + gradients = 44 B/pt (32.8 MB)
    fused_bytes_survey = TRAIN_BATCH * P * (28 + 16 + 16) + TRAIN_BATCH * TRAIN_MAX_GT * 24
    fused_bytes = TRAIN_BATCH * P * (28 + 16) + TRAIN_BATCH * TRAIN_MAX_GT * 24
    focal_bytes = 2 * TRAIN_BATCH * P * NCLS * 4
    fused_traffic, fused_traffic_src = ncu_traffic("assign_loss_fused")
    assign_traffic, assign_traffic_src = ncu_traffic("assign_targets_kernel")
    gbs = lambda nbytes, t_us: nbytes / (t_us * 1e-6) / 1e9               # noqa: E731
    config3 = {
        "workload": f"target assign + GIoU loss fwd+bwd, COCO 832x1344, B={TRAIN_BATCH}, M<={TRAIN_MAX_GT}, P={P}",
        "metric": "target-assign+GIoU us/batch", "value": us, "unit": "us/batch", "min": us_min,
        "path": "FCOSTargetLoss.box_cnt_losses + loss.backward() (autograd), CUDA graph of 4 steps, 5 event pairs",
        "roofline": {"kernel": "b200det_assign_loss_fused (targets + GIoU loss + gradients)", "bound": "hbm",
                     "achieved": gbs(fused_bytes_survey, us_kernels), "peak": peak, "unit": "GB/s",
                     "frac": gbs(fused_bytes_survey, us_kernels) / peak, "bytes_per_launch": fused_bytes_survey,
                     "bytes_note": "SURVEY 8(d) algorithmic bytes (44.7 MB: targets out, reg preds in, reg grads out)",
                     "frac_of_bytes_moved": gbs(fused_bytes, us_kernels) / peak, "bytes_moved": fused_bytes,
                     "us_per_launch": us_kernels, "us_min": us_kernels_min, "traffic": fused_traffic,
                     "traffic_source": fused_traffic_src, "event_pairs": 5},
        "assign": {"kernel": "assign_targets_kernel (FCOSGenTargets.forward)", "bound": "hbm",
                   "achieved": gbs(assign_bytes, us_assign), "peak": peak, "unit": "GB/s",
                   "frac": gbs(assign_bytes, us_assign) / peak, "bytes_per_launch": assign_bytes,
                   "us_per_launch": us_assign, "us_min": us_assign_min, "traffic": assign_traffic,
                   "traffic_source": assign_traffic_src, "event_pairs": 5},
        "with_centerness_us": us_cnt, "unfused_modules_us": us_unfused,
        "full_step": {"what": "targets + focal + centerness + GIoU losses and all gradients (FCOSTargetLoss forward + "
                              "total.backward())", "us_per_batch": us_full, "focal_step_kernel_us": us_focal,
                      "focal_bytes": focal_bytes, "focal_frac_of_peak": gbs(focal_bytes, us_focal) / peak,
                      "focal_step_kernel_fp16_logits_us": us_focal_h}}

    # ---- config 4: dense crowd --------------------------------------------------------------------------
    cand = [W.crowd_candidates(CROWD_N, NCLS, seed=4000 + i) for i in range(CROWD_B)]
    cb = torch.stack([c[0] for c in cand]).to(dev)
    cs = torch.stack([c[1] for c in cand]).to(dev)
    cc = torch.stack([c[2] for c in cand]).to(dev)
    cand1k = [W.crowd_candidates(1000, NCLS, seed=4100 + i) for i in range(BATCH)]
    kb = torch.stack([c[0] for c in cand1k]).to(dev)
    ks = torch.stack([c[1] for c in cand1k]).to(dev)
    kc = torch.stack([c[2] for c in cand1k]).to(dev)
    gt4, lab4 = W.gt_boxes(CROWD_B, CROWD_GT, W.COCO_HW, NCLS, seed=4200)
    gt4, lab4 = gt4.to(dev), lab4.to(dev)
    with sampler:
        us_nms5k, _ = timed_graph(lambda i: ops.batched_nms(cb, cs, cc, SCORE_THR, NMS_THR), 40)
        us_nms1k, _ = timed_graph(lambda i: ops.batched_nms(kb, ks, kc, SCORE_THR, NMS_THR), 80)
        us_assign300, _ = timed_graph(
            lambda i: ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt4, lab4), reps)
    kept = ops.batched_nms(cb, cs, cc, SCORE_THR, NMS_THR)[4]
    config4 = {"workload": f"dense crowd: {CROWD_N} candidates/image x {CROWD_B} images (per-class branch), 1000 crowded "
                           f"candidates x {BATCH} images (coordinate-trick branch), IoU {NMS_THR}; assign with "
                           f"{CROWD_GT} GT x {CROWD_B} images",
               "nms_5000_us_per_batch": us_nms5k, "nms_1000_crowded_us_per_batch": us_nms1k,
               "assign_300gt_us_per_batch": us_assign300, "kept_per_image_5000": [int(v) for v in kept.tolist()],
               "bound": "latency (greedy serial dependency): time only"}
    return config3, config4


def strong_scaling_leg(B, ops, sharding, dist, dev, rank, world, barrier, max_over_ranks, sampler):
    """BASELINE config 5: a batch of 256 split contiguously over the ranks (256/128/64/32 images per GPU).  Per step
    every rank runs the post-process of its shard and one FCOSTargetLoss training step (targets + focal + centerness +
    GIoU losses + all gradients); the packed detections and the rank's share of the batch-mean losses (loss.py:210-213)
    travel to rank 0 in ONE buffer per step (train.py:185-186 gathers its loss the same way)."""
    if STRONG_BATCH % world:
        return None
    lo, hi = sharding.shard_bounds(STRONG_BATCH, world, rank)
    nb = hi - lo
    g_dev = torch.Generator(device=dev).manual_seed(5000 + rank)
    cls = [(torch.randn(nb, NCLS, h, w, device=dev, generator=g_dev) - 4.595).requires_grad_(True) for h, w in W.COCO_LEVELS]
    cnt = [torch.randn(nb, 1, h, w, device=dev, generator=g_dev).requires_grad_(True) for h, w in W.COCO_LEVELS]
    reg = [torch.exp(torch.randn(nb, 4, h, w, device=dev, generator=g_dev) + 3.0).requires_grad_(True) for h, w in W.COCO_LEVELS]
    gt, labels = W.gt_boxes(STRONG_BATCH, TRAIN_MAX_GT, W.COCO_HW, NCLS, seed=5100)
    gt, labels = gt[lo:hi].to(dev), labels[lo:hi].to(dev)
    head = B.FCOSHead(SCORE_THR, NMS_THR, MAX_BOX, W.STRIDES)
    step = B.FCOSTargetLoss(W.STRIDES, W.HISFCOS_RANGES, "giou")
    k_out = min(MAX_BOX, P)
    # the shard is post-processed in micro-batches of 16 forked over up to 8 streams (K1 of one overlaps the
    # per-image select/NMS CTAs of the others, as in the weak-scaling loop); their packed outputs share ONE buffer
    spans = [(i, min(i + BATCH, nb)) for i in range(0, nb, BATCH)]
    sizes = [ops.packed_nbytes(hi_ - lo_, k_out) for lo_, hi_ in spans]
    # ONE buffer per rank for everything that travels: the packed detections of its micro-batches and, in a 256-byte
    # tail, the sums of its per-image losses (cls, cnt, reg) — so the step's only communication is one gather
    pk_all = torch.zeros((sum(sizes) + 256,), dtype=torch.uint8, device=dev)
    pk = pk_all[:sum(sizes)]
    loss_tail = pk_all[sum(sizes):sum(sizes) + 12].view(torch.float32)
    full = torch.empty((world, pk_all.numel()), dtype=torch.uint8, device=dev) if world > 1 else None
    offs = [sum(sizes[:i]) for i in range(len(sizes))]
    # the gather: to rank 0 (where an evaluation / the logging lives, train.py:185-186) by peer-memory pushes on the
    # copy engines + a device barrier (sharding.PeerGather), NCCL all_gather where symmetric memory is unavailable
    peer = None
    if dist is not None and os.environ.get("B200DET_BENCH_GATHER", "peer_root").startswith("peer"):
        try:
            peer = sharding.PeerGather(pk_all.numel(), 2, dev)
        except Exception as e:                              # noqa: BLE001
            print(f"[bench] rank {rank}: config 5 falls back to NCCL ({type(e).__name__}: {e})", file=sys.stderr)
    if dist is not None:
        flag = torch.tensor([1.0 if peer is not None else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag[0]) == 0.0:
            peer = None
    gathers = [0]
    x_mb = [[[t.detach()[lo_:hi_] for t in part] for part in (cls, cnt, reg)] for lo_, hi_ in spans]
    post_streams = [torch.cuda.Stream(device=dev) for _ in range(min(8, len(spans)))]
    side = torch.cuda.Stream(device=dev)

    def post(_i=0):
        cur = torch.cuda.current_stream()
        for st in post_streams:
            st.wait_stream(cur)
        for j, x_j in enumerate(x_mb):
            with torch.cuda.stream(post_streams[j % len(post_streams)]):
                head.detect(x_j, clip_hw=W.COCO_HW, out_packed=pk[offs[j]:offs[j] + sizes[j]])
        for st in post_streams:
            cur.wait_stream(st)

    def train(_i=0):
        for t in cls + cnt + reg:
            t.grad = None
        step([(cls, cnt, reg), gt, labels])[3].backward()
        per = step.per_image                                # this rank's share of the batch-mean losses (loss.py:210-213)
        loss_tail.copy_(torch.stack([per[k].sum() for k in ("cls", "cnt", "reg")]) / STRONG_BATCH)

    def both(_i=0):                                         # the two halves are independent: fork / join
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            post()
        train()
        cur.wait_stream(side)

    def collectives(which):
        """One gather per step: detections and loss sums in one buffer.  Returns the receiver's [world, bytes] view."""
        if dist is None:
            return pk_all[None]
        if peer is not None:
            gathers[0] += 1
            return peer.gather(gathers[0] % 2, pk_all, root=0)
        dist.all_gather_into_tensor(full, pk_all)
        return full

    def timed(fn, which, iters=20, warm=3):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            run, how = g.replay, "CUDA graph of the compute + eager gather"
        except Exception:                                   # noqa: BLE001
            torch.cuda.synchronize()
            run, how = fn, "eager"
        for _ in range(warm):
            run()
            collectives(which)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler:
            e0.record()
            for _ in range(iters):
                run()
                last = collectives(which)
            e1.record()
            barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / iters, how, last

    ms_post, how, _ = timed(post, "post")
    ms_train, _, _ = timed(train, "train")
    ms_both, _, view = timed(both, "both")
    torch.cuda.synchronize()
    n_pk = sum(sizes)
    losses = view[:, n_pk:n_pk + 12].contiguous().view(torch.float32).reshape(-1, 3).sum(dim=0) if rank == 0 else None
    gather_kind = "none" if dist is None else (
        "ONE gather to rank 0 per step (packed detections + the three loss sums in one buffer) by peer-memory pushes + "
        "device barrier" if peer is not None else "ONE NCCL all_gather_into_tensor per step (packed detections + loss sums)")
    return {"workload": f"batch {STRONG_BATCH} split over {world} GPU(s) = {nb} images per GPU: FCOSHead.detect + FCOSTargetLoss "
                        f"forward/backward (M<={TRAIN_MAX_GT}) + gather of the detections and the batch-mean losses",
            "scaling": "strong", "images_per_gpu": nb, "ms_per_step": ms_both, "img_per_s": STRONG_BATCH / (ms_both * 1e-3),
            "postprocess_only_ms": ms_post, "postprocess_img_per_s": STRONG_BATCH / (ms_post * 1e-3),
            "train_step_only_ms": ms_train, "how": how, "iters": 20,
            "batch_mean_losses_cls_cnt_reg": [float(v) for v in losses] if losses is not None else None,
            "collectives": gather_kind,
            "l2": f"per-GPU inputs {nb * 7.94:.0f} MB per step (> 126 MB L2)"}


def eager_reference_leg(dev):
    """The reference's own torch ops (oracle port, which follows head.py / loss.py line by line) run EAGERLY on this
    GPU through stock torch + torchvision CUDA NMS: the practical same-device bar."""
    try:
        import torchvision
        from oracle import fcos_oracle as O
        x = W.head_outputs(BATCH, NCLS, W.COCO_LEVELS, seed=2000)
        x = [[t.to(dev) for t in part] for part in x]
        nms = torchvision.ops.batched_nms

        def post():
            with torch.no_grad():
                dets = O.detect(x, SCORE_THR, NMS_THR, MAX_BOX, W.STRIDES, nms_fn=nms)
                for d in dets:
                    O.clip_boxes_(d[2], *W.COCO_HW)
            return dets

        gt, labels = W.gt_boxes(TRAIN_BATCH, TRAIN_MAX_GT, W.COCO_HW, NCLS, seed=3000)
        gt, labels = gt.to(dev), labels.to(dev)
        regs = [torch.exp(torch.randn(TRAIN_BATCH, 4, h, w, device=dev) + 3).requires_grad_(True) for h, w in W.COCO_LEVELS]

        def train():
            for t in regs:
                t.grad = None
            tgt = O.assign_targets(W.COCO_LEVELS, gt, labels, W.STRIDES, W.HISFCOS_RANGES)
            mask = (tgt[1] > -1).squeeze(-1)
            O.reg_loss(regs, tgt[2], mask, "giou").mean().backward()

        def wall(fn, n):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(n):
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            return min(ts)

        t_post = wall(post, 5)
        t_train = wall(train, 3)
        return {"postprocess_img_per_s": BATCH / t_post, "postprocess_ms_per_batch16": 1e3 * t_post,
                "assign_giou_us_per_batch32": 1e6 * t_train,
                "what": "oracle port (the reference's torch ops in its order) on cuda + torchvision.ops.batched_nms CUDA, eager, "
                        "wall clock incl. its host syncs, best of 5 / 3"}
    except Exception as e:                                  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


def main():
    # stdout carries exactly ONE line, the JSON result: libraries that chat on fd 1 (NCCL prints its version
    # there when NCCL_DEBUG is set) are sent to stderr, the result goes to the saved descriptor
    result_out = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=8, help="batches in flight (CUDA streams) in the timed loop")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, result_out)
    else:
        main_b200(args, rank, world, local, result_out)


if __name__ == "__main__":
    main()
