#!/usr/bin/env python
"""Benchmark of the FCOS detection hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the inference post-process (K1 score -> K2 top-k -> K3 NMS + clip) over one
batch of synthetic head outputs at BASELINE config 2 (COCO 832x1344, 80 classes, batch 16 per
GPU).  Prints ONE JSON line:
  value     whole-job img/s with inputs resident in HBM (rotating over input sets > L2)
  e2e       same metric through the public API (FCOSHead.detect) from pinned HOST buffers,
            H2D of the head outputs and D2H of the detections inside the timed region
  roofline  K1 (score_points, the kernel that moves >95 % of the bytes) timed alone with CUDA
            events: algorithmic bytes / launch time / measured HBM peak
  cpu_baseline  the oracle port of the reference's CPU path on this host, bounded sample
  train     BASELINE config 3 (target assign + GIoU fwd/bwd, B=32, M<=100): us/batch + roofline
With --impl reference the oracle port (the reference is pure Python and does not travel to the
GPU box; the port is pinned against it by tests/golden) is timed on the host cores instead.
Multi-GPU (torchrun): batch sharded by rank (weak scaling, 16 images per GPU); the packed detections of
every round of 8 steps are gathered to rank 0 through NVLink peer memory (sharding.PeerGather; NCCL all_gather
where symmetric memory is unavailable), overlapped with the following rounds; time = max over ranks.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# 8 batches in flight + 2 replay streams + NCCL's stream exceed the default 8 hardware work queues; streams that
# share a queue pick up false dependencies (measured at 2 GPUs: 24.2 -> 22.5 us per step with 32 queues)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402

from pytorch_object_detection_b200 import workloads as W  # noqa: E402

BATCH = 16
NCLS = 80
MAX_BOX = 1000
SCORE_THR = 0.05
NMS_THR = 0.6
TRAIN_BATCH = 32
TRAIN_MAX_GT = 100
P = W.num_points(W.COCO_LEVELS)
WORKLOAD = f"COCO 832x1344 FCOS post-process (score+top-k {MAX_BOX}+NMS {NMS_THR}+clip), {NCLS} classes, " \
           f"P={P}, batch {BATCH} per GPU"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
        except Exception:
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
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
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


def cpu_reference_leg(steps, warmup, sample_images):
    """The reference's CPU path (oracle port) on a bounded sample of the same workload."""
    from oracle import fcos_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    x = W.head_outputs(sample_images, NCLS, W.COCO_LEVELS, seed=1000)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        dets = O.detect(x, SCORE_THR, NMS_THR, MAX_BOX, W.STRIDES)
        for d in dets:
            O.clip_boxes_(d[2], *W.COCO_HW)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": sample_images * len(times) / total, "unit": "img/s", "cores": torch.get_num_threads(),
            "kind": "port", "ms_per_step": 1e3 * total / len(times),
            "sample": f"{sample_images} images of the workload per step x {len(times)} steps, oracle port "
                      f"(torch CPU ops + numpy NMS), {torch.get_num_threads()} torch threads"}


def main():
    # stdout carries exactly ONE line, the JSON result: libraries that chat on fd 1 (NCCL prints its version
    # there when NCCL_DEBUG is set) are sent to stderr, the result goes to the saved descriptor
    result_out = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sets", type=int, default=8, help="input sets rotated through (each 127 MB; L2 is 126 MB)")
    ap.add_argument("--streams", type=int, default=8, help="batches in flight (CUDA streams) in the timed loop; "
                    "measured on B200: 3 -> 26.0, 4 -> 22.6, 6 -> 21.5, 8 -> 21.0 us per batch-16 step (K1 alone: 20.7)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(args.steps, 20)
        leg = cpu_reference_leg(steps, min(warmup, 2), 4)
        line = {"impl": "reference", "metric": "postprocess_throughput", "value": leg["value"], "unit": "img/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(warmup, 2), "ms_per_step": leg["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": WORKLOAD, "sample": leg["sample"]},
                "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": leg["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        result_out.write(json.dumps(line) + "\n")
        result_out.flush()
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU: b200det has no CPU path"
    import pytorch_object_detection_b200 as B
    from pytorch_object_detection_b200 import ops
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

    # ---- synthetic inputs: `sets` independent batches, rotated so every step misses L2 -------
    # (two sets are made on the host — the e2e leg copies them from pinned memory every step —, the others with
    # the same distributions directly on the device: 1 GB of host random numbers per rank is only start-up time)
    host_sets = [W.head_outputs(BATCH, NCLS, W.COCO_LEVELS, seed=2000 + 17 * rank + s) for s in range(min(2, args.sets))]
    dev_sets = [[[t.to(dev) for t in part] for part in hs] for hs in host_sets]
    for s_i in range(len(host_sets), args.sets):
        g_dev = torch.Generator(device=dev).manual_seed(2000 + 17 * rank + s_i)
        cls_d, cnt_d, reg_d = [], [], []
        for h, w in W.COCO_LEVELS:
            cls_d.append(torch.randn(BATCH, NCLS, h, w, device=dev, generator=g_dev) - 4.595)
            cnt_d.append(torch.randn(BATCH, 1, h, w, device=dev, generator=g_dev))
            reg_d.append(torch.exp(torch.randn(BATCH, 4, h, w, device=dev, generator=g_dev) + 3.0))
        dev_sets.append([cls_d, cnt_d, reg_d])
    in_bytes = sum(t.numel() * 4 for part in host_sets[0] for t in part)
    head = B.FCOSHead(SCORE_THR, NMS_THR, MAX_BOX, W.STRIDES)
    sampler = ClockSampler(local)

    # A "round" = one pass over the `sets` input batches = `sets` steps, captured as ONE CUDA graph so the
    # timed loop is not bound by Python launch overhead.  Batches are independent, so inside the graph they
    # are forked round-robin onto `streams` capture streams: K1 (HBM-bound, all SMs) of one batch overlaps
    # the per-image select/NMS kernel (one CTA per image, latency-bound) of the others.
    # Multi-GPU: the path's only collective is the final gather of the detections.  The packed outputs
    # (scores, boxes, classes, keep indices, counts in one allocation) of a whole round sit in one buffer,
    # so ONE NCCL all_gather serves `sets` steps; four output buffers rotate over two streams so that a
    # gather overlaps the kernels of the following rounds and never delays the reuse of its buffer.
    k_out = min(MAX_BOX, P)
    pk_bytes = ops.packed_nbytes(BATCH, k_out)
    for hs in dev_sets:                                   # warm the allocator / library before capture
        head.detect(hs, clip_hw=W.COCO_HW)
    torch.cuda.synchronize()

    def capture_round(n_streams, count=None):
        count = args.sets if count is None else count
        out_big = torch.empty((count, pk_bytes), dtype=torch.uint8, device=dev)
        side = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cs = torch.cuda.current_stream()
            for st in side:
                st.wait_stream(cs)                        # fork
            for s_i, hs in enumerate(dev_sets[:count]):
                with torch.cuda.stream(side[s_i % n_streams]):
                    head.detect(hs, clip_hw=W.COCO_HW, out_packed=out_big[s_i])
            for st in side:
                cs.wait_stream(st)                        # join
        full = torch.empty((world, count * pk_bytes), dtype=torch.uint8, device=dev) if world > 1 and peer is None else None
        return g, out_big, full

    # The gather: peer-memory pushes on the copy engines + a device-side barrier (sharding.PeerGather); NCCL
    # all_gather when symmetric memory is not available (B200DET_BENCH_GATHER=nccl forces it).  The detections
    # are gathered to rank 0, where an evaluation collects them (Test_coco.py:144-168 writes one result file);
    # B200DET_BENCH_GATHER=peer_all gives every rank every detection instead (8 GPUs: 24.2 us per step).
    peer = None
    gather_check = {"ok": None}                             # peer gather compared with NCCL all_gather once
    gather_kind = "none"
    if dist is not None:
        gather_kind = "NCCL all_gather"
        gather_mode = os.environ.get("B200DET_BENCH_GATHER", "peer_root")      # peer_root | peer_all | nccl
        gather_root = 0 if gather_mode == "peer_root" else None
        if gather_mode.startswith("peer"):
            try:
                from pytorch_object_detection_b200.sharding import PeerGather
                peer = PeerGather(args.sets * pk_bytes, 4, dev)
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

    def timed_rounds(n_streams):
        n_buf = 4                                                  # output buffers (graphs) in rotation
        rounds = [list(capture_round(n_streams)) for _ in range(n_buf)]
        n_rounds, tail = divmod(args.steps, args.sets)             # EXACTLY args.steps steps are timed
        if tail:
            rounds.append(list(capture_round(n_streams, tail)))
        pending = [None] * (n_buf + 1)
        # rounds A and B are replayed on two different streams so that the tail of one round overlaps the
        # head of the next (a replay of A still waits for the previous replay of A: same stream)
        outer = [torch.cuda.Stream(device=dev) for _ in range(2 if n_streams > 1 else 1)]
        done = [None] * (n_buf + 1)
        gather_stream = torch.cuda.Stream(device=dev) if dist is not None else None

        def run_round(r, which=None):
            q = r % n_buf if which is None else which
            g, out_big, full = rounds[q]
            with torch.cuda.stream(outer[q % len(outer)]):
                if pending[q] is not None:
                    torch.cuda.current_stream().wait_event(pending[q])   # last gather of this buffer finished
                g.replay()
                done[q] = torch.cuda.Event()
                done[q].record()
            if dist is not None and not os.environ.get("B200DET_BENCH_NO_GATHER"):         # (debug switch)
                # the collective is issued from its own stream, which waits for this round only
                with torch.cuda.stream(gather_stream):
                    gather_stream.wait_event(done[q])
                    if peer is not None and out_big.numel() == peer.nbytes:
                        peer.gather(q % peer.slots, out_big.reshape(-1), root=gather_root)
                    else:
                        if full is None:                    # the shorter tail round
                            full = rounds[q][2] = torch.empty((world, out_big.numel()), dtype=torch.uint8, device=dev)
                        dist.all_gather_into_tensor(full, out_big.reshape(-1))
                    pending[q] = torch.cuda.Event()
                    pending[q].record()

        def drain():
            cur = torch.cuda.current_stream()
            for q in range(n_buf + 1):
                if done[q] is not None:
                    cur.wait_event(done[q])
                if pending[q] is not None:
                    cur.wait_event(pending[q])
                    pending[q] = None

        for r in range(max(n_buf, -(-warmup // args.sets))):
            run_round(r)
        drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler:
            e0.record()
            for st in outer:
                st.wait_event(e0)                         # nothing starts before the start event
            t_host = time.perf_counter()
            for r in range(n_rounds):
                run_round(r)
            if tail:
                run_round(0, which=n_buf)
            t_host = time.perf_counter() - t_host
            if os.environ.get("B200DET_BENCH_TRACE"):
                print(f"[bench] rank {rank}: host issue time {1e6 * t_host / max(1, args.steps):.2f} us/step "
                      f"({n_streams} streams)", file=sys.stderr)
            drain()                                       # the last gathers are inside the timed region
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
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
        return ms / args.steps, args.steps

    # ---- value: device-resident, `n_streams` batches in flight; and one batch at a time ------------
    n_streams = max(1, min(args.streams, args.sets))
    ms_per_step, steps_timed = timed_rounds(n_streams)
    value = world * BATCH / (ms_per_step * 1e-3)
    ms_single, _ = timed_rounds(1)
    launches = steps_timed * 2                            # score_points + fused select/NMS kernel per step

    # ---- e2e: pinned host inputs -> H2D -> public API -> D2H of the detections -----------------
    # Two staging slots: the H2D copies of step i+1 (split over two copy streams) overlap the kernels and
    # the D2H of step i; the host waits for every step's packed result in pinned memory.
    pinned = [[[t.pin_memory() for t in part] for part in hs] for hs in host_sets[:2]]
    e2e_steps = max(3, min(args.steps, 30))
    d2h_bytes = pk_bytes
    slots = []
    for k in range(2):
        stage = [[torch.empty_like(t, device=dev) for t in part] for part in host_sets[0]]
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

    def e2e_issue(i):
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

    def e2e_run(n):
        for i in range(n):
            e2e_issue(i)
            if i >= 1:
                slots[(i - 1) % 2]["done"].synchronize()   # host has step i-1's result
        slots[(n - 1) % 2]["done"].synchronize()

    e2e_run(3)
    barrier()
    with sampler:
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        barrier()
        e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = world * BATCH * e2e_steps / e2e_s

    # ---- roofline of the dominant kernel (K1) timed alone, same inputs, same rotation ----------
    peak, peak_src = peaks()
    k1_bytes = BATCH * P * ((NCLS + 1) * 4 + 4 + 2)       # cls + cnt planes read, score f32 + class i16 written
    # K1 is launched through the same C-ABI entry the fused call uses.  One CUDA graph holds `chunk`
    # back-to-back launches rotating over the input sets, so the CUDA events bracket kernel time, not
    # Python or graph-launch latency.
    for hs in dev_sets:
        ops.score_points(hs[0], hs[1], W.STRIDES)
    torch.cuda.synchronize()
    chunk = 2 * args.sets                                   # launches per graph replay / event pair
    k1_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(k1_graph):
        for i in range(chunk):
            hs = dev_sets[i % args.sets]
            ops.score_points(hs[0], hs[1], W.STRIDES)
    for i in range(max(1, warmup // chunk + 1)):
        k1_graph.replay()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(max(1, args.steps // chunk))]
    with sampler:
        for a, b in evs:
            a.record()
            k1_graph.replay()
            b.record()
        torch.cuda.synchronize()
    k1_ms = sorted(a.elapsed_time(b) / chunk for a, b in evs)
    k1_avg = sum(k1_ms) / len(k1_ms)
    achieved = k1_bytes / (k1_avg * 1e-3) / 1e9
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this size from the ncu
    # --set full capture in profiles/ (120.63 MB read + 3.5..6.4 MB written per launch)
    roofline = {"kernel": "score_points_kernel (K1)", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": 124.9e6, "peak_source": peak_src,
                "bytes_per_launch": k1_bytes, "us_per_launch": 1e3 * k1_avg, "us_median": 1e3 * k1_ms[len(k1_ms) // 2]}

    # ---- training side (config 3): target assignment + GIoU fwd/bwd ---------------------------
    train = None
    if rank == 0:
        gt, labels = W.gt_boxes(TRAIN_BATCH, TRAIN_MAX_GT, W.COCO_HW, NCLS, seed=3000)
        gt, labels = gt.to(dev), labels.to(dev)
        regs = [[torch.exp(torch.randn(TRAIN_BATCH, 4, h, w, device=dev) + 3).requires_grad_(True)
                 for h, w in W.COCO_LEVELS] for _ in range(4)]
        gen = B.FCOSGenTargets(W.STRIDES, W.HISFCOS_RANGES)
        fake = [torch.empty(TRAIN_BATCH, 1, h, w, device="meta") for h, w in W.COCO_LEVELS]

        cnts = [[torch.randn(TRAIN_BATCH, 1, h, w, device=dev).requires_grad_(True) for h, w in W.COCO_LEVELS]
                for _ in range(4)]
        fused = B.FCOSTargetLoss(W.STRIDES, W.HISFCOS_RANGES, "giou")

        def train_step(i):
            """Fused path: targets + GIoU loss + its gradients (3 PDL-chained launches) behind autograd."""
            for t in regs[i % 4]:
                t.grad = None                              # optimizer.zero_grad(set_to_none=True)
            loss, _ = fused.box_cnt_losses(None, regs[i % 4], gt, labels)
            loss.backward()
            return loss

        def train_step_cnt(i):
            """Same with the centerness BCE branch (loss.py:29-57) in the same launches."""
            for t in regs[i % 4] + cnts[i % 4]:
                t.grad = None
            reg_loss, cnt_loss = fused.box_cnt_losses(cnts[i % 4], regs[i % 4], gt, labels)
            (reg_loss + cnt_loss).backward()

        def train_step_unfused(i):
            """The drop-in modules one by one: FCOSGenTargets -> compute_reg_loss -> backward (3 kernels + glue)."""
            for t in regs[i % 4]:
                t.grad = None
            tgt = gen([[fake, fake, fake], gt, labels])
            loss = B.compute_reg_loss(regs[i % 4], tgt[2], None, "giou", _mask_src=tgt[1]).mean()
            loss.backward()
            return loss

        clss = []            # two sets of class logits (238 MB each), made when the full step is timed

        def train_step_full(i):
            """The whole FCOSGenTargets + FCOSLoss step: fused assign/box/centerness launch + one focal launch that
            writes the loss and the 238 MB class gradient from one read of the logits."""
            for t in regs[i % 4] + cnts[i % 4] + clss[i % 2]:
                t.grad = None
            fused([(clss[i % 2], cnts[i % 4], regs[i % 4]), gt, labels])[3].backward()

        def train_step_full_two_focal_kernels(i):
            """Same step with the focal loss as separate forward and backward kernels (logits read twice)."""
            for t in regs[i % 4] + cnts[i % 4] + clss[i % 2]:
                t.grad = None
            reg_loss, cnt_loss = fused.box_cnt_losses(cnts[i % 4], regs[i % 4], gt, labels)
            cls_t, cnt_t, _ = fused.targets
            cls_loss = B.compute_cls_loss(clss[i % 2], cls_t, None, _mask_src=cnt_t).mean()
            (cls_loss + cnt_loss + reg_loss).backward()

        def timed_graph(fn, reps):
            """us per call of fn(i), captured as one CUDA graph of 4 calls (no Python between launches)."""
            for i in range(4):
                fn(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = max(1, reps // 4)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(4):
                        fn(i)
                run = g.replay
            except Exception:                               # autograd inside capture refused: time eagerly
                torch.cuda.synchronize()
                run = lambda: [fn(i) for i in range(4)]
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            a.record()
            for _ in range(n):
                run()
            b.record()
            torch.cuda.synchronize()
            return 1e3 * a.elapsed_time(b) / (4 * n)

        with sampler:
            us = timed_graph(train_step, args.steps)
            us_cnt = timed_graph(train_step_cnt, args.steps)
            us_unfused = timed_graph(train_step_unfused, args.steps)
            us_kernels = timed_graph(lambda i: ops.assign_loss_fused(regs[i % 4], None, W.STRIDES, W.HISFCOS_RANGES,
                                                                    gt, labels, 1), args.steps)
            clss.extend([(torch.randn(TRAIN_BATCH, NCLS, h, w, device=dev) - 4.595).requires_grad_(True)
                          for h, w in W.COCO_LEVELS] for _ in range(2))
            us_full = timed_graph(train_step_full, args.steps)
            us_full_two = timed_graph(train_step_full_two_focal_kernels, args.steps)
            us_focal = timed_graph(lambda i: ops.cls_loss_step(clss[i % 2], fused.targets[0],
                                                               num_pos=fused.per_image["num_pos"]), args.steps)
            # the autocast case (train.py:175): fp16 class logits read as they are, fp16 gradients written
            cls_h = [[t.detach().half() for t in clss[i]] for i in range(2)]
            scale_state = torch.tensor([65536.0, 0.0], device=dev)
            us_focal_h = timed_graph(lambda i: ops.cls_loss_step(cls_h[i % 2], fused.targets[0],
                                                                 num_pos=fused.per_image["num_pos"],
                                                                 up_mean=scale_state), args.steps)
            del cls_h
            # assign alone, for its own roofline: 28 bytes written per point
            us_assign = timed_graph(
                lambda i: ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt, labels), args.steps)
        assign_bytes = TRAIN_BATCH * P * 28 + TRAIN_BATCH * TRAIN_MAX_GT * 24
        fused_bytes = TRAIN_BATCH * P * (28 + 16) + TRAIN_BATCH * TRAIN_MAX_GT * 24   # targets + reg gradients written
        train = {"workload": f"target assign + GIoU loss fwd+bwd, COCO 832x1344, B={TRAIN_BATCH}, M<={TRAIN_MAX_GT}",
                 "us_per_batch": us, "path": "FCOSTargetLoss (fused: count + tile + finalize kernels, PDL-chained) + autograd",
                 "fused_kernels_us": us_kernels, "fused_bytes": fused_bytes,
                 "fused_gbs": fused_bytes / (us_kernels * 1e-6) / 1e9,
                 "fused_frac_of_peak": fused_bytes / (us_kernels * 1e-6) / 1e9 / peak,
                 "with_centerness_us": us_cnt, "unfused_us_per_batch": us_unfused,
                 "full_step": {"what": "targets + focal + centerness + GIoU losses and all gradients (FCOSTargetLoss "
                                       "forward + total.backward())",
                               "us_per_batch": us_full, "us_with_two_focal_kernels": us_full_two,
                               "focal_step_kernel_us": us_focal, "focal_bytes": 2 * TRAIN_BATCH * P * NCLS * 4,
                               "focal_gbs": 2 * TRAIN_BATCH * P * NCLS * 4 / (us_focal * 1e-6) / 1e9,
                               "focal_frac_of_peak": 2 * TRAIN_BATCH * P * NCLS * 4 / (us_focal * 1e-6) / 1e9 / peak,
                               "focal_step_kernel_fp16_logits_us": us_focal_h},
                 "assign_us": us_assign, "assign_bytes": assign_bytes,
                 "assign_gbs": assign_bytes / (us_assign * 1e-6) / 1e9,
                 "assign_frac_of_peak": assign_bytes / (us_assign * 1e-6) / 1e9 / peak}

    clocks = sampler.summary()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cpu = cpu_reference_leg(steps=6, warmup=1, sample_images=4)
    line = {"metric": "postprocess_throughput", "value": value, "unit": "img/s", "n_gpus": world, "steps": steps_timed,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * BATCH,
                       "l2": f"{args.sets} input sets of {in_bytes / 1e6:.0f} MB rotated (> 126 MB L2)",
                       "in_flight": f"{n_streams} batches on {n_streams} CUDA streams inside one CUDA graph of {args.sets} steps",
                       "collective": f"one gather of the packed detections per {args.sets} steps, overlapped: {gather_kind}" if world > 1 else "none",
                       "gather_equals_nccl_all_gather": gather_check["ok"]},
            "single_stream": {"value": world * BATCH / (ms_single * 1e-3), "unit": "img/s", "ms_per_step": ms_single},
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "steps": e2e_steps, "api": "FCOSHead.detect(clip_hw=...) on pinned host inputs"},
            "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "train": train, "clocks": clocks}
    result_out.write(json.dumps(line) + "\n")
    result_out.flush()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
