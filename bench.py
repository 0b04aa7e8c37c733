#!/usr/bin/env python
"""bench.py -- x-vector extraction frames/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path

A *step* is one pass of the extraction hot path (pack -> 5 fused tcgen05 TDNN layers, the last one pooling in its epilogue ->
pool statistics -> embed_layer-0) over one batch of synthetic MFCC: BASELINE.json configs[1], 256 utterances x 400
frames x 23 ceps per GPU (weak scaling: every rank gets its own batch; for N > 1,
like the product, ONE NCCL gather of all embeddings to rank 0 ends the job, inside the timed region:
the only collective on the path).

  value  frames/s with the features already resident in HBM (CUDA events around each step on the
         launching stream, L2 flushed between steps, max over ranks).
  e2e    frames/s through the reference-facing C-ABI call (the sess.run boundary, reference
         local/tf/models.py:412-415) in its pipelined form xv_submit_host / xv_collect: pinned HOST
         features in, HOST embeddings out, every step's H2D + D2H inside the timed region, two
         steps in flight so the copy of one overlaps the kernels of the previous one.
  roofline      the fused TDNN layer kernel (5 launches/step) against the measured tensor peak;
                per-launch figures for all launches in roofline.launches.
  cpu_baseline  the reference's CPU path restated in torch fp32 (oracle/xvector_torch_cpu.py --
                TensorFlow 1.x cannot be installed here), run the way the reference runs it: one
                utterance per call, 2 threads per process, cores/2 processes
                (models.py:361-363,410-414; extract_xvectors.sh:63,83).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "xvector_extraction_frames_per_sec"
UNIT = "frames/s"
FEAT_DIM = 23
EMB_DIM = 512
L2_FLUSH_BYTES = 512 << 20            # > 126 MB L2

TOPOLOGIES = {   # reference local/tf/models.py:443-445 and :545-548
    "ModelWithoutDropoutTdnn": dict(kernel_sizes=[5, 3, 3, 1, 1], dilations=[1, 2, 3, 1, 1],
                                    layer_sizes=[512, 512, 512, 512, 1536], embedding_sizes=[512, 512]),
    "ModelWithoutDropout": dict(kernel_sizes=[5, 5, 7, 1, 1], dilations=[1, 1, 1, 1, 1],
                                layer_sizes=[512, 512, 512, 512, 1536], embedding_sizes=[512, 512]),
    # nonlinearity variants of the dense topology (models.py:643-744 per-channel PReLU, :866-983 leaky_relu 0.2)
    "ModelWithoutDropoutPRelu": dict(kernel_sizes=[5, 5, 7, 1, 1], dilations=[1, 1, 1, 1, 1],
                                     layer_sizes=[512, 512, 512, 512, 1536], embedding_sizes=[512, 512], act="prelu"),
    "ModelL2LossWithoutDropoutLRelu": dict(kernel_sizes=[5, 5, 7, 1, 1], dilations=[1, 1, 1, 1, 1],
                                           layer_sizes=[512, 512, 512, 512, 1536], embedding_sizes=[512, 512], act="lrelu"),
    # self-attention pooling (models.py:990-1051)
    "ModelL2LossWithoutDropoutLReluAttention": dict(kernel_sizes=[5, 5, 7, 1, 1], dilations=[1, 1, 1, 1, 1],
                                                    layer_sizes=[512, 512, 512, 512, 3072], embedding_sizes=[512, 512],
                                                    act="lrelu", pooling="attention"),
}


def flop_per_frame(topo):
    """Algorithmic FLOP per frame of each frame layer (2 x taps x C_in x C_out)."""
    out, prev = [], FEAT_DIM
    for k, w in zip(topo["kernel_sizes"], topo["layer_sizes"]):
        out.append(2 * k * prev * w)
        prev = w
    return out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), tflops=float(p["bf16_tflops"]),
                    tflops_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback")


NCU_TRAFFIC_FILE = "r02_ncu_full_summary.json"      # ncu --set full of the SHIPPED extraction kernels at configs[1], both topologies


def ncu_traffic_bytes(topology):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the fused TDNN layer kernel, averaged over the
    frame-layer launches of ONE captured extraction step of `topology` at configs[1], from the committed summary
    profiles/r02_ncu_full_summary.json (written by tools/ncu_summary.py from the capture of tools/profile_step.py).
    (None, None) if that file or that topology's capture is absent -- never another workload's numbers."""
    path = os.path.join(ROOT, "profiles", NCU_TRAFFIC_FILE)
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        rows = [r for r in json.load(f) if r.get("report", "").endswith("prof_%s.ncu-rep" % topology) and
                ("tdnn_pair_kernel<0" in r.get("kernel", "") or "tdnn_pair_kernel<1" in r.get("kernel", "") or
                 "tdnn_tail_fused_kernel" in r.get("kernel", "") or "tdnn_first_kernel" in r.get("kernel", ""))]
    vals = [(r["dram_read_MB"] + r["dram_write_MB"]) * 1e6 for r in rows if r.get("dram_read_MB") is not None]
    return (round(sum(vals) / len(vals)) if vals else None), NCU_TRAFFIC_FILE


# ------------------------------------------------------------------------------------ clocks
BAD_REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")


class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons of one GPU through NVML while a region runs."""
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
            0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    _nvml = {}      # cuda index -> (pynvml, handle, sm_max) or an error string; filled ONCE per process (nvmlInit and the
                    # handle lookup serialise across the processes of a node: r1 paid them inside the timed window)

    @classmethod
    def prepare(cls, cuda_index):
        if cuda_index in cls._nvml:
            return
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            cls._nvml[cuda_index] = (pynvml, h, int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
        except Exception as e:      # clocks are evidence, not the product: report why they are missing
            cls._nvml[cuda_index] = repr(e)

    def __init__(self, cuda_index, period=0.004):
        super().__init__(daemon=True)
        self.period = period
        self.samples, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self._stop_evt = threading.Event()
        self.ok = False
        self.prepare(cuda_index)
        st = self._nvml[cuda_index]
        if isinstance(st, str):
            self.err = st
        else:
            self.nv, self.h, self.sm_max = st
            self.ok = True

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = int(get_reasons(self.h))
                for bit, name in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join()
        if not self.ok or not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.sm_max, reasons=[], note=getattr(self, "err", "no samples"))
        out = dict(sm_mhz=int(statistics.median(self.samples)), sm_max_mhz=self.sm_max,
                   reasons=sorted(self.reasons), samples=len(self.samples))
        if self.power:
            out.update(power_w_median=round(statistics.median(self.power), 1), power_w_max=round(max(self.power), 1))
        return out


def clocks_rejected(c):
    if any(r in BAD_REASONS for r in c.get("reasons", [])):
        return True
    if c.get("sm_mhz") and c.get("sm_max_mhz") and not c.get("reasons") and c["sm_mhz"] < 0.6 * c["sm_max_mhz"]:
        return True        # stuck low with no reason: a leftover clock lock
    return False


# ------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from xvector_b200 import _native, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs one process per GPU: launch with python -m torch.distributed.run "
                             "--nnodes=1 --nproc-per-node %d --master-addr 127.0.0.1 bench.py --gpus %d ..." %
                             (args.gpus, args.gpus, args.gpus))
        raise SystemExit("WORLD_SIZE=%d does not match --gpus %d" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py measures the sm_100a path: no CUDA device is visible (there is no CPU fallback; "
                         "use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    topo = TOPOLOGIES[args.topology]
    params = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"],
                                   weight_set=args.weight_set, activation=topo.get("act", "relu"),
                                   pooling=topo.get("pooling", "stats"))
    eng = _native.XvecEngine(topo["kernel_sizes"], topo["dilations"], topo["layer_sizes"], EMB_DIM, FEAT_DIM,
                             device=local_rank, activation=topo.get("act", "relu"), pooling=topo.get("pooling", "stats"))
    eng.set_params(params)
    for kv in args.option:
        eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))

    ClockSampler.prepare(local_rank)               # nvmlInit + handle lookup OUTSIDE every timed window
    B, T = args.batch, args.frames
    lens = np.full(B, T, np.int32)
    frames = int(lens.sum())
    feats_np = synthetic.mfcc_batch(2 + 1000 * rank, lens)             # configs[1]: seed 2 (rank 0)
    feats_host = torch.empty((frames, FEAT_DIM), dtype=torch.float32, pin_memory=True)
    feats_host.numpy()[:] = feats_np
    emb_host = torch.empty((B, EMB_DIM), dtype=torch.float32, pin_memory=True)
    feats_dev = feats_host.to(dev)
    emb_dev = torch.empty((B, EMB_DIM), dtype=torch.float32, device=dev)
    # N > 1: rank 0 owns the job's result table [N x steps x B, 512] in its HBM, every rank maps it (CUDA IPC) and the
    # kernel that finishes a batch (utt_average_kernel) stores its rows straight into the table over NVLink: the
    # "gather to rank 0" is that store, there is no end-of-job collective (r1 had one dist.gather, whose window soaked
    # up 12 ms of rank skew at N = 8).  A barrier after the last step orders the writers before rank 0's read.
    peer = None
    if world > 1:
        table_rows = world * args.steps * B
        if rank == 0:
            peer = _native.PeerTable.create(local_rank, table_rows, EMB_DIM)
        box = [peer.handle if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            peer = _native.PeerTable.open(local_rank, table_rows, EMB_DIM, box[0])
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def out_of(i):
        """Where step i's x-vectors go: the local buffer (N = 1, warm-up) or this rank's rows of rank 0's table."""
        if peer is None or i is None:
            return emb_dev
        return peer.data_ptr((rank * args.steps + i) * B)

    def step_resident(i=None):
        # the product's per-batch call: forward + make_embedding's chunk average (one chunk per utterance here)
        eng.forward_utts(feats_dev, lens, out_of(i), stream=stream)

    feats_host2 = [feats_host, feats_host.clone().pin_memory()]
    emb_host2 = [emb_host, torch.empty_like(emb_host).pin_memory()]

    def run_e2e(steps):
        """K steps through xv_submit_host_utts / xv_collect (pinned host in, pinned host out), two in flight:
        the H2D copy of step i+1 overlaps the kernels of step i.  Returns wall seconds."""
        timed = steps == args.steps
        t0 = time.perf_counter()
        prev = None
        for i in range(steps):
            slot = i & 1
            ticket = eng.submit_host_utts(feats_host2[slot], lens, out_host=emb_host2[slot],
                                          out_dev=(out_of(i) if (timed and peer is not None) else None))
            if prev is not None:
                eng.collect(prev)
            prev = ticket
        eng.collect(prev)
        if world > 1 and timed:
            dist.barrier()                                             # every rank's rows are in rank 0's table
            if rank == 0:
                peer.read(peer.rows - B, B)                          # rank 0 reads the job's result (last rows: rank N-1's)
        return time.perf_counter() - t0

    def timed_resident(steps):
        sampler = ClockSampler(local_rank)
        barrier(); torch.cuda.synchronize(dev)
        w0 = time.perf_counter()
        sampler.start()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i, (s, e) in enumerate(evs):
            flush.zero_()                                              # evict L2 between steps (untimed)
            s.record(stream)
            step_resident(i)
            e.record(stream)
        torch.cuda.synchronize(dev)
        w1 = time.perf_counter()
        barrier()                                                      # N > 1: all writers done -> the table is complete
        w2 = time.perf_counter()
        clocks = sampler.finish()
        ms = [s.elapsed_time(e) for s, e in evs]
        # own_ms: this rank's wall time for its K steps incl. the untimed L2 flushes; wait_ms: what it then waited for the
        # slowest rank at the closing barrier (rank skew + barrier latency; there is no data left to move)
        return ms, clocks, dict(own_ms=(w1 - w0) * 1e3, wait_ms=(w2 - w1) * 1e3)

    def timed_e2e(steps):
        barrier(); torch.cuda.synchronize(dev)
        sec = run_e2e(steps)
        torch.cuda.synchronize(dev); barrier()
        return sec * 1e3

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(x):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = x
        dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    # ---- warm-up, then the timed regions ------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    if peer is not None:                          # first touch of the peer mapping
        step_resident(0)
    torch.cuda.synchronize(dev)
    barrier()
    eng.check_overflow()
    ms, clocks, walls = timed_resident(args.steps)
    remeasured = False
    if clocks_rejected(clocks):
        remeasured = True
        ms, clocks, walls = timed_resident(args.steps)
    per_rank_ms = all_ranks(sum(ms))
    total_ms = max(per_rank_ms)
    ms_per_step = total_ms / args.steps
    value = world * frames / (ms_per_step * 1e-3)
    multi = None
    if world > 1:
        # the job's result as rank 0 sees it: every row of the table written (finite, non-zero), rank 0's own rows
        # bit-identical to a local run of the same batch
        ok = None
        if rank == 0:
            got = peer.read()
            step_resident()
            torch.cuda.synchronize(dev)
            mine = emb_dev.cpu().numpy()
            ok = bool(np.isfinite(got).all() and (np.abs(got).max(axis=1) > 0).all()
                      and all(np.array_equal(got[i * B:(i + 1) * B], mine) for i in range(args.steps)))
        multi = dict(gather="none: utt_average_kernel stores each batch's rows into rank 0's table over NVLink (CUDA IPC peer "
                            "mapping); one barrier after the last step",
                     step_ms_sum_per_rank=[round(v, 4) for v in per_rank_ms],
                     skew_ms=round(max(per_rank_ms) - min(per_rank_ms), 4),
                     own_wall_ms_per_rank=[round(v, 3) for v in all_ranks(walls["own_ms"])],
                     closing_barrier_wait_ms_per_rank=[round(v, 3) for v in all_ranks(walls["wait_ms"])],
                     table_rows=peer.rows, table_complete_and_rank0_rows_bit_identical=ok)

    run_e2e(max(args.warmup, 3))
    e2e_total = max_over_ranks(timed_e2e(args.steps))
    e2e_value = world * frames * args.steps / (e2e_total * 1e-3)

    # ---- configs[2]-like ragged batch (200-1000 frames per utterance), device-resident, same frame budget ----
    rag_lens = synthetic.lengths_uniform(3, 4096)
    rag_lens = rag_lens[:int(np.searchsorted(np.cumsum(rag_lens), frames))].astype(np.int32)
    rag_feats = torch.from_numpy(synthetic.mfcc_batch(3 + 1000 * rank, rag_lens)).to(dev)
    rag_emb = torch.empty((len(rag_lens), EMB_DIM), dtype=torch.float32, device=dev)
    for _ in range(3):
        eng.forward(rag_feats, rag_lens, emb_dev=rag_emb, stream=stream)
    # ragged and uniform calls ALTERNATE inside one window: this block runs late in a long, power-limited process, and
    # a drifting clock would otherwise be read as a cost of raggedness (profiles/r01_ragged_vs_uniform.txt)
    rag_evs, uni_evs = [], []
    for _ in range(min(args.steps, 50)):
        for evs_, feats_, lens_, emb_ in ((rag_evs, rag_feats, rag_lens, rag_emb), (uni_evs, feats_dev, lens, emb_dev)):
            flush.zero_()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record(stream); eng.forward(feats_, lens_, emb_dev=emb_, stream=stream); e_.record(stream)
            evs_.append((s_, e_))
    torch.cuda.synchronize(dev)
    rag_ms = float(np.mean([s_.elapsed_time(e_) for s_, e_ in rag_evs]))
    uni_ms = float(np.mean([s_.elapsed_time(e_) for s_, e_ in uni_evs]))
    ragged = dict(workload="configs[2]-like: %d utterances of 200-1000 frames in one call (packed rows, no bucketing needed)"
                           % len(rag_lens), frames_per_step_per_gpu=int(rag_lens.sum()), ms_per_step=round(rag_ms, 5),
                  value_per_gpu=round(float(rag_lens.sum()) / (rag_ms * 1e-3), 1), unit=UNIT,
                  control=dict(note="configs[1]'s uniform batch, calls alternating with the ragged ones in the same window",
                               ms_per_step=round(uni_ms, 5), value_per_gpu=round(frames / (uni_ms * 1e-3), 1)))

    # ---- per-launch durations (CUDA events on the launching stream, inside the library) -------
    eng.set_option("profile", 1)
    per_launch = []
    for _ in range(3):
        step_resident()
    for _ in range(min(args.steps, 50)):
        flush.zero_()
        eng.forward_utts(feats_dev, lens, emb_dev, stream=stream)
        per_launch.append(eng.last_kernel_ms())
    eng.set_option("profile", 0)
    launches_per_step = eng.last_launch_count
    kms = np.asarray(per_launch, dtype=np.float64).mean(axis=0)        # [7]
    peaks = load_peaks()
    fl = flop_per_frame(topo)
    c_last = topo["layer_sizes"][-1]
    k0_pad = -(-topo["kernel_sizes"][0] * FEAT_DIM // 128) * 128
    n_blk = B * (-(-T // 32))
    # (kernel, bound, algorithmic FLOP or bytes per launch) in launch order
    n_tail = 2 + (3 if topo.get("pooling") == "attention" else 0)       # launches behind the frame layers (+ attention's three)
    # launches in front of pool_stats: [pack +] layer 0 (+ splice), layers 1..n-3, then the last two layers as one or two launches
    n_front = len(kms) - 1 - n_tail
    # len(fl) - 1: splice inside layer 0's kernel AND the last two layers as one launch (the default topologies);
    # len(fl) + 1: neither (attention pooling / split precision); len(fl): read as "pack launch + fused last two layers"
    fused_first = n_front == len(fl) - 1
    fused_tail = n_front <= len(fl)
    # layer 0 (K = 128) is bound by its own output stream (K0 = 117 760 FLOP / (92 + 1 024) B = 105 FLOP/B, HBM-bound):
    # judged against HBM with SURVEY's algorithmic bytes per frame (fp32 features in, fp16 activations out); layers 1.. tensor bound
    if fused_first:
        table = [("tdnn_first_kernel[splice+L0]", "hbm", frames * (FEAT_DIM * 4 + topo["layer_sizes"][0] * 2))]
    else:
        table = [("pack_im2col_kernel", "hbm", frames * (FEAT_DIM * 4 + k0_pad * 2 + 1)),
                 ("tdnn_pair_kernel[L0]", "hbm", frames * (FEAT_DIM * 4 + topo["layer_sizes"][0] * 2))]
    if fused_tail:
        table += [("tdnn_pair_kernel[L%d]" % i, "tensor", frames * f) for i, f in enumerate(fl) if 0 < i < len(fl) - 2]
        table += [("tdnn_tail_fused_kernel[L%d+L%d]" % (len(fl) - 2, len(fl) - 1), "tensor", frames * (fl[-2] + fl[-1]))]
    else:
        table += [("tdnn_pair_kernel[L%d]" % i, "tensor", frames * f) for i, f in enumerate(fl) if i > 0]
    if topo.get("pooling") == "attention":       # models.py:1037-1051: score GEMM [frames, C] x [C, C], softmax over time, weighted sums
        c_last //= 2
        table += [("tdnn_pair_kernel<3>[attention scores]", "tensor", frames * 2 * c_last * c_last),
                  ("attn_softmax_kernel", "hbm", frames * (2 * (c_last // 256) + 2) * 4),
                  ("attn_pool_kernel", "hbm", frames * c_last * 2 + n_blk * 2 * c_last * 4)]
    table += [("pool_stats_kernel", "hbm", n_blk * 2 * c_last * 4 + B * 2 * c_last * 6)]
    # embed_layer-0 on tensor cores: split-fp16 GEMM (3 products), then K-split reduction + make_embedding's chunk average
    table += [("tdnn_pair_kernel<2>[embed_layer-0]", "tensor", 2 * B * 2 * c_last * EMB_DIM),
              ("embed_reduce_utt_kernel", "hbm", B * EMB_DIM * 4 * 2 + B * 16)]
    launches = []
    for (name, bound, work), ms in zip(table, kms):
        d = dict(kernel=name, ms=round(float(ms), 5), bound=bound)
        if bound == "tensor":
            tf = work / (ms * 1e-3) / 1e12
            d.update(achieved=round(tf, 1), unit="TFLOP/s", frac=round(tf / peaks["tflops"], 4))
        else:
            gbs = work / (ms * 1e-3) / 1e9
            d.update(achieved=round(gbs, 1), unit="GB/s", frac=round(gbs / peaks["hbm_gbs"], 4))
        launches.append(d)
    n_layer_launches = len(fl) - 1 if fused_tail else len(fl)
    first_layer = 0 if fused_first else 1                               # index of layer 0's launch
    tdnn_ms = float(kms[first_layer:first_layer + n_layer_launches].sum())
    tdnn_tf = frames * sum(fl) / (tdnn_ms * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic_bytes(args.topology)
    roofline = dict(kernel="%stdnn_pair_kernel x%d%s (%d launches/step, figures are per-step sums / averages)"
                           % ("tdnn_first_kernel + " if fused_first else "",
                              n_layer_launches - (1 if fused_tail else 0) - (1 if fused_first else 0),
                              " + tdnn_tail_fused_kernel" if fused_tail else "", n_layer_launches),
                    bound="tensor", achieved=round(tdnn_tf, 1), peak=peaks["tflops"], unit="TFLOP/s",
                    frac=round(tdnn_tf / peaks["tflops"], 4), traffic=traffic,
                    traffic_source=("profiles/%s: mean DRAM bytes per launch over the %d frame-layer launches of one captured step of %s "
                                    "at configs[1]" % (traffic_src, len(fl), args.topology)) if traffic is not None else None,
                    flop_per_launch_avg=round(frames * sum(fl) / n_layer_launches),
                    peak_source="%s bf16/fp16 burst (MEASURED_PEAKS.json)" % peaks["source"],
                    share_of_step=round(tdnn_ms / float(kms.sum()), 4),
                    step_frac_of_tensor_peak=round(value / world * sum(fl) / 1e12 / peaks["tflops"], 4),
                    step_frac_of_sustained_peak=round(value / world * sum(fl) / 1e12 / peaks["tflops_sustained"], 4),
                    peak_sustained=peaks["tflops_sustained"],
                    launches=launches)

    # ---- sustained regime: >= 2 s of back-to-back steps (what a 1 M-utterance job lives in: the power cap settles) --------
    def timed_run(engine, seconds):
        engine.forward_utts(feats_dev, lens, emb_dev, stream=stream)
        torch.cuda.synchronize(dev)
        n = max(200, int(seconds / (ms_per_step * 1e-3)))
        sampler = ClockSampler(local_rank, period=0.02)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.start()
        e0.record(stream)
        for _ in range(n):
            engine.forward_utts(feats_dev, lens, emb_dev, stream=stream)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n, n, sampler.finish()

    sus_ms, sus_n, sus_clocks = timed_run(eng, 2.0)
    sustained = dict(note="back-to-back steps, no L2 flush between them (the layer activations, 100+ MB each, exceed L2 anyway)",
                     steps=sus_n, seconds=round(sus_ms * sus_n * 1e-3, 3), ms_per_step=round(sus_ms, 5),
                     value_per_gpu=round(frames / (sus_ms * 1e-3), 1), unit=UNIT,
                     frac_of_sustained_tensor_peak=round(frames * sum(fl) / (sus_ms * 1e-3) / 1e12 / peaks["tflops_sustained"], 4),
                     frac_of_burst_tensor_peak=round(frames * sum(fl) / (sus_ms * 1e-3) / 1e12 / peaks["tflops"], 4),
                     peak_sustained=peaks["tflops_sustained"], clocks=sus_clocks)

    # ---- the recipe's default topology (run_xvector.sh:90 ModelWithoutDropout: dense kernels 5,5,7,1,1) -----------------
    dense = None
    if args.topology != "ModelWithoutDropout":
        dtopo = TOPOLOGIES["ModelWithoutDropout"]
        deng = _native.XvecEngine(dtopo["kernel_sizes"], dtopo["dilations"], dtopo["layer_sizes"], EMB_DIM, FEAT_DIM, device=local_rank)
        deng.set_params(synthetic.make_params(dtopo["kernel_sizes"], dtopo["layer_sizes"], dtopo["embedding_sizes"],
                                              weight_set=args.weight_set))
        for _ in range(3):
            deng.forward_utts(feats_dev, lens, emb_dev, stream=stream)
        devs = []
        for _ in range(min(args.steps, 50)):
            flush.zero_()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record(stream); deng.forward_utts(feats_dev, lens, emb_dev, stream=stream); e_.record(stream)
            devs.append((s_, e_))
        torch.cuda.synchronize(dev)
        d_ms = float(np.mean([s_.elapsed_time(e_) for s_, e_ in devs]))
        dfl = flop_per_frame(dtopo)
        deng.set_option("profile", 1)
        dk = []
        for _ in range(10):
            flush.zero_()
            deng.forward_utts(feats_dev, lens, emb_dev, stream=stream)
            dk.append(deng.last_kernel_ms())
        deng.set_option("profile", 0)
        dk = np.asarray(dk, dtype=np.float64).mean(axis=0)
        d_n = len(dfl) - 1                                                 # last two layers fused: one launch fewer
        d_first = len(dk) - 3 - d_n                                        # 1 with a separate pack launch, 0 with the splice in layer 0's kernel
        d_layers = float(dk[d_first:d_first + d_n].sum())
        d_tf = frames * sum(dfl) / (d_layers * 1e-3) / 1e12
        dense = dict(workload="configs[1] with ModelWithoutDropout (taps %s)" % dtopo["kernel_sizes"], ms_per_step=round(d_ms, 5),
                     value_per_gpu=round(frames / (d_ms * 1e-3), 1), unit=UNIT,
                     roofline=dict(bound="tensor", achieved=round(d_tf, 1), peak=peaks["tflops"], unit="TFLOP/s",
                                   frac=round(d_tf / peaks["tflops"], 4), share_of_step=round(d_layers / float(dk.sum()), 4),
                                   traffic=ncu_traffic_bytes("ModelWithoutDropout")[0],
                                   layer_ms=[round(float(v), 5) for v in dk[d_first:d_first + d_n]]),
                     step_frac_of_tensor_peak=round(frames * sum(dfl) / (d_ms * 1e-3) / 1e12 / peaks["tflops"], 4))
        deng.close()

    frontend = measure_frontend(args, eng, dev, stream, flush, rank, peaks)

    out = dict(metric=METRIC, value=round(value, 1), unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
               ms_per_step=round(ms_per_step, 5), higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f16", data="synthetic",
               config=dict(workload="configs[1]: batch=%d utterances x %d frames x %d-dim MFCC per GPU, %s (%s), "
                                    "weight set %s" % (B, T, FEAT_DIM, args.topology,
                                                       "taps %s dil %s" % (topo["kernel_sizes"], topo["dilations"]),
                                                       args.weight_set),
                           frames_per_step_per_gpu=frames, l2="flushed between steps (%d MiB write)" % (L2_FLUSH_BYTES >> 20),
                           parallelism=("utterance sharding x%d; each batch's x-vectors are stored into rank 0's result table over "
                                        "NVLink peer memory by the kernel that produces them (no gather collective; NCCL only for "
                                        "barriers)" % world) if world > 1 else "single GPU",
                           arithmetic="fp16 operands (RN), fp32 accumulate (tcgen05 kind::f16), fp32 epilogue/pooling"),
               e2e=dict(value=round(e2e_value, 1), unit=UNIT, h2d_bytes_per_step=frames * FEAT_DIM * 4 + B * 3 * 4,
                        d2h_bytes_per_step=B * EMB_DIM * 4 + 4, ms_per_step=round(e2e_total / args.steps, 5),
                        api="xv_submit_host_utts / xv_collect, 2 in flight (pinned host buffers in and out"
                            + ("; rows also stored to rank 0's table, closing barrier + rank 0's read inside the window)" if world > 1 else ")")),
               gpu_launches=int(launches_per_step * args.steps),
               clocks=clocks, roofline=roofline, sustained=sustained, dense_topology=dense, ragged=ragged, frontend=frontend)
    if multi is not None:
        out["multi_gpu"] = multi
    if remeasured:
        out["clocks"]["note"] = "first measurement rejected (throttle reason / low clocks); this is the re-measurement"

    if peer is not None:
        torch.cuda.synchronize(dev)
        barrier()
        peer.close()
    eng.close()
    if not args.no_product:
        out["product"] = measure_product(args, dev, rank, world, params, topo)
        out["config4_synthetic_job"] = measure_config4(args, dev, rank, world, params, topo)
    if not args.no_train and topo.get("act", "relu") == "relu" and topo.get("pooling", "stats") == "stats":
        out["train_step"] = measure_train_step(args, dev, rank, world, peaks, topo)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["reader"] = measure_reader()
        out["cpu_baseline"] = cpu_baseline_subprocess(args)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


def measure_frontend(args, eng, dev, stream, flush, rank, peaks):
    """The Kaldi pipe of local/tf/extract_xvectors.sh:68 on the device (include/xvec_frontend.h): sliding-window CMVN
    (window 300, centred) + voiced-frame selection over configs[1]'s geometry taken as RAW frames with a synthetic VAD
    track; kernel time by CUDA events on the launching stream (L2 flushed between iterations), judged against HBM with
    the algorithmic bytes (raw rows + VAD in, voiced rows out), and the raw host path (xv_submit_host_raw) end to end."""
    import torch
    from xvector_b200 import _native, synthetic
    B, T = args.batch, args.frames
    lens = np.full(B, T, np.int32)
    raw = synthetic.mfcc_batch(7 + 1000 * rank, lens) + np.float32(-30.0) * (np.arange(FEAT_DIM) == 0).astype(np.float32)
    rng = np.random.default_rng(7 + rank)
    tracks = [synthetic.synthetic_vad(rng, T) for _ in range(B)]
    for v in tracks:
        if v.sum() < 25:                      # every utterance passes the extractor's min-chunk-size rule
            v[:min(T, 100)] = 1.0
    vad = np.concatenate(tracks)
    keep = vad.reshape(B, T).astype(bool).sum(axis=1).astype(np.int32)
    raw_host = torch.from_numpy(raw).pin_memory()
    vad_host = torch.from_numpy(vad).pin_memory()
    raw_dev, vad_dev = raw_host.to(dev), vad_host.to(dev)
    out_dev = torch.empty((int(keep.sum()), FEAT_DIM), dtype=torch.float32, device=dev)
    opts = _native.XvCmvnOpts()
    for _ in range(3):
        eng.frontend(raw_dev, vad_dev, lens, keep, opts, out_dev=out_dev, stream=stream)
    evs = []
    n_it = min(args.steps, 50)
    for _ in range(n_it):
        flush.zero_()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record(stream); eng.frontend(raw_dev, vad_dev, lens, keep, opts, out_dev=out_dev, stream=stream); e_.record(stream)
        evs.append((s_, e_))
    torch.cuda.synchronize(dev)
    ms = float(np.mean([s_.elapsed_time(e_) for s_, e_ in evs]))
    alg_bytes = int(B * T * (FEAT_DIM * 4 + 4) + int(keep.sum()) * FEAT_DIM * 4)
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    # the same kernel on a problem large enough for an HBM fraction to mean something: 10 240 utterances x 400 raw frames
    # (4.1 M raw frames, 0.38 GB in) -- the batch above tiled 40 times
    large = None
    try:
        rep = 40
        lens_l, keep_l = np.tile(lens, rep), np.tile(keep, rep)
        raw_l, vad_l = raw_dev.repeat(rep, 1), vad_dev.repeat(rep)
        out_l = torch.empty((int(keep_l.sum()), FEAT_DIM), dtype=torch.float32, device=dev)
        for _ in range(2):
            eng.frontend(raw_l, vad_l, lens_l, keep_l, opts, out_dev=out_l, stream=stream)
        evl = []
        for _ in range(10):
            flush.zero_()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record(stream); eng.frontend(raw_l, vad_l, lens_l, keep_l, opts, out_dev=out_l, stream=stream); e_.record(stream)
            evl.append((s_, e_))
        torch.cuda.synchronize(dev)
        ms_l = float(np.mean([s_.elapsed_time(e_) for s_, e_ in evl]))
        gbs_l = alg_bytes * rep / (ms_l * 1e-3) / 1e9
        large = dict(workload="%d x %d RAW frames (%.1f M)" % (B * rep, T, B * rep * T / 1e6), ms_per_call=round(ms_l, 5),
                     raw_frames_per_sec=round(B * rep * T / (ms_l * 1e-3), 1),
                     roofline=dict(bound="hbm", achieved=round(gbs_l, 1), peak=peaks["hbm_gbs"], unit="GB/s",
                                   frac=round(gbs_l / peaks["hbm_gbs"], 4), algorithmic_bytes_per_call=alg_bytes * rep))
        del raw_l, vad_l, out_l
    except Exception as err:                                     # noqa: BLE001
        large = dict(error="%s: %s" % (type(err).__name__, err))
    # raw host path: pinned raw rows + VAD in, embeddings out, two submissions in flight
    emb = [torch.empty((B, EMB_DIM), dtype=torch.float32).pin_memory() for _ in range(2)]
    def run(n):
        pending = None
        for i in range(n):
            t = eng.submit_host_raw(raw_host, vad_host, lens, keep, keep, emb[i % 2], opts)
            if pending is not None:
                eng.collect(pending)
            pending = t
        eng.collect(pending)
    run(3)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    run(n_it)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_it
    # cpu_baseline leg (rank 0, N = 1 only): the plain-C restatement of the Kaldi pipe, one core, bounded sample
    cpu = None
    if rank == 0 and int(os.environ.get("WORLD_SIZE", "1")) == 1 and not args.no_cpu_baseline:
        cpu = frontend_cpu_baseline(raw, vad, B, T)
    return dict(workload="apply-cmvn-sliding(300, centred) | select-voiced-frames on %d x %d RAW frames, %.0f %% voiced"
                         % (B, T, 100.0 * float(keep.sum()) / (B * T)),
                kernels="cmvn_select_kernel (one launch; vad_tile_count_kernel in front only for utterances > 4096 frames)", ms_per_call=round(ms, 5),
                raw_frames_per_sec=round(B * T / (ms * 1e-3), 1),
                roofline=dict(bound="hbm", achieved=round(gbs, 1), peak=peaks["hbm_gbs"], unit="GB/s",
                              frac=round(gbs / peaks["hbm_gbs"], 4), algorithmic_bytes_per_call=alg_bytes),
                e2e=dict(api="xv_submit_host_raw / xv_collect, 2 in flight (raw rows + VAD track from pinned host memory, "
                             "front end + network on the device, embeddings back)",
                         ms_per_step=round(e2e_ms, 5), raw_frames_per_sec=round(B * T / (e2e_ms * 1e-3), 1),
                         voiced_frames_per_sec=round(float(keep.sum()) / (e2e_ms * 1e-3), 1),
                         h2d_bytes_per_step=int(B * T * (FEAT_DIM * 4 + 4) + B * 4 * 3), d2h_bytes_per_step=B * EMB_DIM * 4 + 4),
                large_problem=large, cpu_baseline=cpu)


def _scratch_dir(need_bytes=16 << 30):
    """Scratch for the product-path blocks: /dev/shm when it has room (the input is meant to sit in the page cache), else the
    default temporary directory."""
    import shutil
    import tempfile
    base = None
    try:
        if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) and shutil.disk_usage("/dev/shm").free > need_bytes:
            base = "/dev/shm"
    except OSError:
        base = None
    return tempfile.mkdtemp(prefix="xvec_product_", dir=base)


def write_model_dir(path, topology, params, num_classes=0):
    """A model directory Model.load_model reads (model.meta + model.npz + done) holding the bench's synthetic weights."""
    from xvector_b200 import models
    topo = TOPOLOGIES[topology]
    meta = dict(pooling=topo.get("pooling", "stats"), format=models.META_FORMAT, model_class=topology, num_classes=int(num_classes),
                input_feature_dim=FEAT_DIM, kernel_sizes=list(topo["kernel_sizes"]), dilation_rates=list(topo["dilations"]),
                layer_sizes=list(topo["layer_sizes"]), embedding_sizes=list(topo["embedding_sizes"]), activation=topo.get("act", "relu"))
    models.Model.save_model(models._Session(params, meta), path, None)


def measure_product(args, dev, rank, world, params, topo):
    """BASELINE configs[2] (N = 1) / configs[3] scaled to 12 500 utterances per GPU (N > 1) through the PRODUCT entry point:
    Model.make_embedding from a feature ark FILE (200-1000 frames per utterance, seed 3) to an x-vector ark + scp written by
    rank 0 -- wall clock around the call on every rank, barrier on both sides, max over ranks.  Inside the window: header
    index of the rank's byte stripe, the stripe-boundary exchange, pread into page-locked batches, H2D, the network, the
    chunk average, the stores into rank 0's table (N > 1), the closing barrier, rank 0's read of the table, formatting
    and writing the output files.  Outside: writing the synthetic input (page cache), loading the model on the first call."""
    import shutil
    import torch
    import torch.distributed as dist
    from xvector_b200 import kaldi_io, models, synthetic
    out = dict()
    tmp = None
    try:
        n_per_rank = 10000 if world == 1 else 12500
        box = [None]
        if rank == 0:
            tmp = _scratch_dir()
            box[0] = tmp
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        tmp = box[0]
        # every rank writes its share of the input (the same 1250-utterance block of synthetic MFCC under fresh keys: content
        # does not matter to throughput, and 100 k utterances need not be generated), rank 0 concatenates
        lens = synthetic.lengths_uniform(3, 1250 if world > 1 else n_per_rank)
        feats = synthetic.mfcc_batch(3, lens)
        offs = np.concatenate([[0], np.cumsum(lens)])
        part = os.path.join(tmp, "part.%d" % rank)
        with open(part, "wb") as f:
            for i in range(n_per_rank):
                j = i % len(lens)
                kaldi_io.write_mat(f, feats[offs[j]:offs[j + 1]], key="utt%07d" % (rank * n_per_rank + i))
        n_frames_rank = int(lens.sum()) * (n_per_rank // len(lens)) + int(lens[:n_per_rank % len(lens)].sum())
        path = os.path.join(tmp, "feats.ark")
        if world > 1:
            dist.barrier()
        if rank == 0:
            with open(path, "wb") as dst:
                for r in range(world):
                    with open(os.path.join(tmp, "part.%d" % r), "rb") as src:
                        shutil.copyfileobj(src, dst, 64 << 20)
                    os.unlink(os.path.join(tmp, "part.%d" % r))
            write_model_dir(os.path.join(tmp, "model"), args.topology, params)
        if world > 1:
            dist.barrier()
        total_frames = n_frames_rank * world
        model = getattr(models, args.topology)()
        runs = []
        for it in range(3):
            ark, scp = os.path.join(tmp, "xvector.%d.ark" % it), os.path.join(tmp, "xvector.%d.scp" % it)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            if rank == 0:
                with kaldi_io.open_vector_writer("ark,scp:%s,%s" % (ark, scp)) as w:
                    model.make_embedding(path, w, os.path.join(tmp, "model"), 25, 10000, True, None)
            else:
                model.make_embedding(path, None, os.path.join(tmp, "model"), 25, 10000, True, None)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            runs.append(dt)
        file_stats = dict(getattr(model, "last_job_stats", {}))
        peer_variant = None
        if world > 1:
            # the same job with rank 0's output a STREAM (what a `| copy-vector ...` wspecifier is): rows go to rank 0's
            # peer-memory table over NVLink and rank 0 writes them all
            import io
            peer_runs = []
            for _ in range(2):           # the first run also pays NCCL's connection set-up for the key gather (a pattern no earlier
                dist.barrier()           # block of this process used); the second is the job itself
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                buf = io.BytesIO() if rank == 0 else None
                model.make_embedding(path, buf, os.path.join(tmp, "model"), 25, 10000, True, None)
                torch.cuda.synchronize(dev)
                dist.barrier()
                dt = time.perf_counter() - t0
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                peer_runs.append(float(t.item()))
            t = torch.tensor([min(peer_runs)], dtype=torch.float64)
            same = None
            if rank == 0:
                with open(os.path.join(tmp, "xvector.2.ark"), "rb") as f:
                    same = f.read() == buf.getvalue()
            peer_variant = dict(output="in-memory stream on rank 0 (peer-memory result table)", seconds=round(float(t.item()), 4),
                                all_runs_s=[round(r, 4) for r in peer_runs],
                                value=round(total_frames / float(t.item()), 1), bytes_identical_to_the_file_output=same,
                                rank0_breakdown_s={k: (round(v, 4) if isinstance(v, float) else v)
                                                   for k, v in getattr(model, "last_job_stats", {}).items()})
        check = None
        if rank == 0:
            # the job's output: every key once and in order, the scp points at the same vectors, sampled rows bit-identical
            # to the device path run directly on those utterances
            ark = os.path.join(tmp, "xvector.2.ark")
            n_out, last, ok_order = 0, None, True
            sample = {}
            want_idx = {0, 1, n_per_rank * world - 1, (n_per_rank * world) // 2}
            for i, (k, v) in enumerate(kaldi_io.read_vec_flt_ark(ark)):
                ok_order = ok_order and k == "utt%07d" % i
                if i in want_idx:
                    sample[i] = v
                n_out += 1
            eng = model._engine
            same = True
            for i, v in sample.items():
                j = (i % n_per_rank) % len(lens)
                x = torch.from_numpy(feats[offs[j]:offs[j + 1]]).to(dev)
                o = torch.empty((1, EMB_DIM), dtype=torch.float32, device=dev)
                eng.forward_utts(x, np.array([lens[j]], np.int32), o)
                torch.cuda.synchronize(dev)
                same = same and np.array_equal(o.cpu().numpy()[0], v)
            first = open(os.path.join(tmp, "xvector.2.scp")).readline().split()
            check = dict(utterances_written=n_out, keys_in_input_order=bool(ok_order), scp_first_line=first[0] if first else None,
                         sampled_rows_bit_identical_to_direct_forward=bool(same), output_bytes=os.path.getsize(ark))
        warm = min(runs[1:])
        out = dict(workload=("configs[2]: %d utterances of 200-1000 frames (%.1f M frames, %.0f MB feature ark file in the page "
                             "cache) -> x-vector ark + scp, Model.make_embedding" if world == 1 else
                             "configs[3] scaled: %d utterances of 200-1000 frames (%.1f M frames, %.0f MB feature ark file in the "
                             "page cache), byte-striped over the ranks -> ONE x-vector ark + scp written by rank 0, "
                             "Model.make_embedding under torchrun") % (n_per_rank * world, total_frames / 1e6,
                                                                         total_frames * 92 / 1e6),
                   unit=UNIT, value=round(total_frames / warm, 1), seconds=round(warm, 4),
                   first_call=dict(seconds=round(runs[0], 4), value=round(total_frames / runs[0], 1),
                                   note="includes Model.load_model, weight upload and the first allocation of page-locked "
                                        "batch buffers / device workspaces"),
                   all_runs_s=[round(r, 4) for r in runs],
                   rank0_breakdown_s={k: (round(v, 4) if isinstance(v, float) else v) for k, v in file_stats.items()},
                   batch_frames=int(os.environ.get("XVEC_BATCH_FRAMES", "0")) or "default", reader_threads=os.environ.get("XVEC_READER_THREADS", "default"),
                   check=check)
        if peer_variant is not None:
            out["stream_output_variant"] = peer_variant
    except Exception as err:                                     # noqa: BLE001  (a diagnostic block must not cost the bench line)
        import traceback
        out = dict(error="%s: %s" % (type(err).__name__, err), trace=traceback.format_exc()[-1500:])
    finally:
        if world > 1:
            try:
                dist.barrier()
            except Exception:                                    # noqa: BLE001
                pass
        if rank == 0 and tmp is not None:
            shutil.rmtree(tmp, ignore_errors=True)
    return out


def measure_config4(args, dev, rank, world, params, topo):
    """BASELINE configs[3]: "1 M-utterance synthetic extraction sharded across 8 x B200, gather to rank 0 ark", scaled to 125 000
    utterances per GPU (N = 8: the full 1 M).  55 GB of features cannot be held or fed by the host (hard part 7), so
    every rank GENERATES its utterances' rows on the device (xv_synth_mfcc, reproducible on the host) and the rest is the
    product's job loop: network, chunk average, every rank writing its byte range of the ONE x-vector ark (+ scp lines to rank 0).
    Wall clock, barrier on both sides, max over ranks; no host->device feature copy is in it, and the line says so."""
    import shutil
    import torch
    import torch.distributed as dist
    from xvector_b200 import ark_job, kaldi_io, models, synthetic
    tmp = None
    try:
        n_total = 125000 * world
        box = [None]
        if rank == 0:
            tmp = _scratch_dir()
            box[0] = tmp
            write_model_dir(os.path.join(tmp, "model"), args.topology, params)
        if world > 1:
            dist.broadcast_object_list(box, src=0)
            dist.barrier()
        tmp = box[0]
        model = getattr(models, args.topology)()
        model.load_model(None, os.path.join(tmp, "model"), None)
        engine = model._get_engine(dev.index)
        batch_frames = int(os.environ.get("XVEC_BATCH_FRAMES", "400000"))
        runs = []
        for it in range(2):
            ark, scp = os.path.join(tmp, "xv4.%d.ark" % it), os.path.join(tmp, "xv4.%d.scp" % it)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            src = ark_job.SyntheticSource(dev.index, n_total, 4, batch_frames)
            counts = src.counts("cuda:%d" % dev.index)
            writer = kaldi_io.open_vector_writer("ark,scp:%s,%s" % (ark, scp)) if rank == 0 else None
            try:
                model._run_extraction_job(src, counts, writer, engine, dev.index, 25, None, time.time(), time.time())
            finally:
                if writer is not None:
                    writer.close()
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            runs.append(dt)
            frames_total = int(counts[:, 3].sum())
        check = None
        if rank == 0:
            # spot checks: the ark holds every key once, in order; sampled rows equal the device path run on the SAME utterance
            # regenerated on the host (numpy restatement of the generator) -- bit for bit
            lens = synthetic.lengths_uniform(4, n_total)
            want_idx = sorted({0, 1, n_total // 2, n_total - 1})
            got, n_out, ordered = {}, 0, True
            for i, (k, v) in enumerate(kaldi_io.read_vec_flt_ark(os.path.join(tmp, "xv4.1.ark"))):
                ordered = ordered and (i % 50021 != 0 or k == "utt%07d" % i)
                if i in want_idx:
                    ordered = ordered and k == "utt%07d" % i
                    got[i] = v
                n_out += 1
            same = True
            for i, v in got.items():
                x = torch.from_numpy(synthetic.counter_mfcc(4, i, int(lens[i]))).to(dev)
                o = torch.empty((1, EMB_DIM), dtype=torch.float32, device=dev)
                engine.forward_utts(x, np.array([lens[i]], np.int32), o)
                torch.cuda.synchronize(dev)
                same = same and np.array_equal(o.cpu().numpy()[0], v)
            check = dict(utterances_written=n_out, keys_in_order=bool(ordered), output_bytes=os.path.getsize(os.path.join(tmp, "xv4.1.ark")),
                         sampled_rows_bit_identical_to_forward_of_host_regenerated_features=bool(same))
        best = min(runs)
        return dict(workload="configs[3] at %d utterances (125 000 per GPU) of 200-1000 frames, %.1f M frames; features generated on the "
                             "device per rank, x-vectors into ONE ark + scp (%s)" % (n_total, frames_total / 1e6,
                                                                                      "each rank writes its byte range" if world > 1 else "one rank"),
                    unit=UNIT, value=round(frames_total / best, 1), seconds=round(best, 4), all_runs_s=[round(r, 4) for r in runs],
                    h2d_feature_bytes=0, note="no host->device feature traffic in this figure: see `product` for the ark-file-fed job",
                    rank0_breakdown_s={k: (round(v, 4) if isinstance(v, float) else v) for k, v in getattr(model, "last_job_stats", {}).items()},
                    check=check)
    except Exception as err:                                     # noqa: BLE001
        import traceback
        return dict(error="%s: %s" % (type(err).__name__, err), trace=traceback.format_exc()[-1500:])
    finally:
        if world > 1:
            try:
                dist.barrier()
            except Exception:                                    # noqa: BLE001
                pass
        if rank == 0 and tmp is not None:
            shutil.rmtree(tmp, ignore_errors=True)


def measure_reader(n_utts=1500, passes=3):
    """Host side of the ark -> x-vector product path, alone (no kernel runs): Model._read_batches over a configs[2]-like ark
    FILE (200-1000 frames per utterance, in the page cache) into the page-locked staging buffers, batches dropped by a null
    consumer; once with the sequential stream parser (one thread) and once with the native header index (xv_ark_scan) +
    pooled pread jobs (the default for regular files).  Never fails the bench: errors are reported in the block."""
    import queue
    import shutil
    import tempfile
    from xvector_b200 import kaldi_io, models, synthetic
    tmp = tempfile.mkdtemp(prefix="xvec_reader_")
    saved = os.environ.get("XVEC_READER_THREADS")
    try:
        path = os.path.join(tmp, "feats.ark")
        lens = synthetic.lengths_uniform(3, n_utts)
        with open(path, "wb") as f:
            for i, n in enumerate(lens):
                kaldi_io.write_mat(f, synthetic.mfcc(1000 + i, int(n)), key="utt%07d" % i)
        out = dict(workload="%d utterances of 200-1000 frames x %d (%.0f MB ark file, page cache) -> staging buffers"
                            % (n_utts, FEAT_DIM, os.path.getsize(path) / 1e6), unit="frames/s")
        for label, threads in (("sequential_parser_1_thread", "1"), ("indexed_pooled_pread_default", None)):
            if threads is None:
                os.environ.pop("XVEC_READER_THREADS", None)
            else:
                os.environ["XVEC_READER_THREADS"] = threads
            model = models.Model.__new__(models.Model)
            staging = models._Staging(FEAT_DIM, 400000)
            work = queue.Queue(maxsize=3)
            counters = dict(total_segments=0, total_segments_len=0, num_fail=0, num_success=0)

            def consume():
                while True:
                    b = work.get()
                    if b is None:
                        return
                    staging.release(b.slot)

            th = threading.Thread(target=consume)
            th.start()
            try:
                with open(path, "rb") as f:                      # warm-up: page cache, staging buffers
                    model._read_batches(f, staging, work, counters, 25, 10000, 400000, 0, 1, None)
                counters["total_segments_len"] = 0
                t0 = time.perf_counter()
                for _ in range(passes):
                    with open(path, "rb") as f:
                        model._read_batches(f, staging, work, counters, 25, 10000, 400000, 0, 1, None)
                dt = time.perf_counter() - t0
            finally:
                work.put(None)
                th.join()
            out[label] = round(counters["total_segments_len"] / dt, 1)
        return out
    except Exception as err:                                     # noqa: BLE001  (a diagnostic block must not cost the bench line)
        return dict(error="%s: %s" % (type(err).__name__, err))
    finally:
        if saved is None:
            os.environ.pop("XVEC_READER_THREADS", None)
        else:
            os.environ["XVEC_READER_THREADS"] = saved
        shutil.rmtree(tmp, ignore_errors=True)


def frontend_cpu_baseline(raw, vad, n_utt, T, seconds=2.0):
    """oracle/kaldi_frontend_oracle.c (gcc -O2, Kaldi's running-sum recursion in double; what apply-cmvn-sliding |
    select-voiced-frames do on a CPU) over the bench's utterances, one core, for about `seconds`."""
    from oracle import kaldi_frontend_c as kfc
    kfc.frontend(raw[:T], vad[:T])                       # build / load / warm up
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        i = done % n_utt
        kfc.frontend(raw[i * T:(i + 1) * T], vad[i * T:(i + 1) * T])
        done += 1
    cpu_s = time.perf_counter() - t0
    return dict(value=round(done * T / cpu_s, 1), unit="raw frames/s", cores=1, kind="port",
                sample="%d utterances of %d frames through oracle/kaldi_frontend_oracle.c (gcc -O2), one core" % (done, T))


def measure_train_step(args, dev, rank, world, peaks, topo):
    """BASELINE configs[4]: one data-parallel training step (TDNN + stats pooling + softmax over 5000 speakers,
    64 x 400 frames per GPU): xv_train_forward_backward -> NCCL all-reduce of the flat gradient (N > 1) -> xv_train_apply.
    Reported as an extra block of the JSON line (the headline metric stays extraction frames/s)."""
    import torch
    import torch.distributed as dist
    from xvector_b200 import _native, synthetic
    B, T, NC, steps, warm = 64, 400, 5000, 60, 5
    P = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], num_classes=NC, weight_set="B")
    eng = _native.XvecEngine(topo["kernel_sizes"], topo["dilations"], topo["layer_sizes"], EMB_DIM, FEAT_DIM, device=dev.index)
    tr = _native.XvecTrainer(eng, NC, topo["embedding_sizes"][1])
    tr.set_params(P)
    feats_host = torch.empty((B * T, FEAT_DIM), dtype=torch.float32, pin_memory=True)
    feats_host.numpy()[:] = synthetic.mfcc(5 + 1000 * rank, B * T)                      # configs[4]: seed 5
    lab_host = torch.from_numpy(np.random.default_rng(5 + rank).integers(0, NC, B).astype(np.int32)).pin_memory()
    feats, lab = feats_host.to(dev), lab_host.to(dev)
    grad = torch.zeros(tr.n_grad, dtype=torch.float32, device=dev) if world > 1 else None
    la_host = torch.zeros(2, dtype=torch.float32, pin_memory=True)
    comm = torch.cuda.Stream(dev) if world > 1 else None

    def step(e2e):
        if e2e:
            feats.copy_(feats_host, non_blocking=True)
            lab.copy_(lab_host, non_blocking=True)
        if world > 1:        # gradient buckets are all-reduced on `comm` as they become final, under the rest of the backward
            la = tr.forward_backward_allreduce(feats, lab, B, T, grad, torch.cuda.current_stream(dev), comm)
        else:
            la = tr.forward_backward(feats, lab, B, T, grad_dev=grad)
        tr.apply(1e-4, grad_dev=grad, grad_scale=1.0 / world)
        if e2e:
            la_host.copy_(la, non_blocking=True)
        return la

    def timed(e2e):
        for _ in range(warm):
            step(e2e)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(e2e)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    ms_res = timed(False)
    eng.check_overflow()
    ms_e2e = timed(True)
    loss = float(la_host[0])
    launches = tr.last_launch_count
    kernels = None
    if world == 1:
        eng.set_option("profile", 1)
        agg = {}
        for _ in range(3):
            step(False)
            torch.cuda.synchronize(dev)
            for n, t_ in zip(tr.last_kernel_names(), eng.last_kernel_ms()):
                agg[n] = agg.get(n, 0.0) + t_ / 3
        eng.set_option("profile", 0)
        kernels = {n: round(v, 4) for n, v in sorted(agg.items(), key=lambda kv: -kv[1])}
    fl = sum(flop_per_frame(topo))
    # forward + data gradient (all layers but the first) + weight gradient, frame level only
    flop_step = B * T * (3 * fl - flop_per_frame(topo)[0])
    tensor_ms = sum(v for n, v in (kernels or {}).items() if "pair_kernel" in n)
    out = dict(workload="configs[4]: training step, %s + stats pooling + softmax over %d speakers, batch %d x %d frames per GPU, "
                        "fp16 operands / fp32 accumulate + fp32 master weights, Adam" % (args.topology, NC, B, T),
               ms_per_step=round(ms_res, 4), value=round(world * B * T / (ms_res * 1e-3), 1), unit="frames/s",
               e2e=dict(ms_per_step=round(ms_e2e, 4), value=round(world * B * T / (ms_e2e * 1e-3), 1), unit="frames/s",
                        h2d_bytes_per_step=B * T * FEAT_DIM * 4 + B * 4, d2h_bytes_per_step=8),
               parallelism="data parallel x%d: NCCL all-reduce of the flat fp32 gradient (%.1f MB) per step in two buckets: the segment-level %.1f MB under the frame-level backward, the frame-level rest behind the step"
                           % (world, tr.n_params * 4 / 1e6, (tr.n_params - tr.seg_grad_offset) * 4 / 1e6)
               if world > 1 else "single GPU",
               gpu_launches_per_step=launches, loss_after=round(loss, 4),
               frame_level_flop_per_step=flop_step,
               step_frac_of_tensor_peak=round(flop_step / (ms_res * 1e-3) / 1e12 / peaks["tflops"], 4))
    if kernels:
        out["kernel_ms"] = kernels
        out["tensor_kernels"] = dict(ms=round(tensor_ms, 4), achieved=round(flop_step / (tensor_ms * 1e-3) / 1e12, 1), unit="TFLOP/s",
                                     frac=round(flop_step / (tensor_ms * 1e-3) / 1e12 / peaks["tflops"], 4))
    tr.close()
    eng.close()
    if world == 1 and not args.no_cpu_baseline:
        # the torch-CPU restatement of the same training step (fp32, all host threads) on a bounded sample: 16 of the 64 segments
        import time
        from oracle import xvector_train_oracle as tro
        Bc = 16
        x = synthetic.mfcc(5, Bc * T).reshape(Bc, T, FEAT_DIM)
        labels = np.random.default_rng(5).integers(0, NC, Bc)
        tro.forward_backward(x[:2], labels[:2], P, topo, dtype=torch.float32)          # warm-up
        t0 = time.time()
        tro.forward_backward(x, labels, P, topo, dtype=torch.float32)
        dt = time.time() - t0
        out["cpu_baseline"] = dict(value=round(Bc * T / dt, 1), unit="frames/s", cores=torch.get_num_threads(), kind="port",
                                   sample="1 step of %d x %d frames (forward + backward, no optimizer), torch fp32 autograd restatement" % (Bc, T))
    return out


def cpu_baseline_subprocess(args):
    """The CPU leg runs in a fresh interpreter (it forks worker processes; this one holds a CUDA context)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
           "--topology", args.topology, "--weight-set", args.weight_set, "--frames", str(args.frames),
           "--step-seconds", "4"]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:
        return dict(value=None, unit=UNIT, cores=None, kind="port", sample="failed: %r" % (e,))


# ------------------------------------------------------------------------------------ reference arm
def _cpu_worker(wid, params, topology, feats, n_utts_max, frames, ready, go, done, counts, threads):
    import torch
    from oracle.xvector_torch_cpu import TorchCpuXvector
    torch.set_num_threads(threads)
    net = TorchCpuXvector(params, topology)
    net.forward(feats[:frames])                                        # warm-up (oneDNN primitive creation)
    while True:
        ready.wait()
        n = go.value
        if n <= 0:
            return
        for u in range(n):
            off = ((wid * 7 + u) % n_utts_max) * frames
            net.forward(feats[off:off + frames])                       # ONE utterance per call (models.py:410-414)
        counts[wid] = n
        done.wait()


def run_reference(args):
    """Times the reference's CPU extraction path.  TensorFlow 1.x is not installable here, so the
    timed code is the torch fp32 restatement under oracle/ ("port"), run in the reference's own
    operating mode: B=1 per call, 2 threads per process, cores/2 processes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from xvector_b200 import synthetic
    topo = TOPOLOGIES[args.topology]
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 2)
    threads = 2
    procs = max(1, min(cores // threads, 64))
    params = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"],
                                   weight_set=args.weight_set, activation=topo.get("act", "relu"),
                                   pooling=topo.get("pooling", "stats"))
    frames = args.frames
    n_utts_max = 16
    feats = synthetic.mfcc_batch(2, np.full(n_utts_max, frames, np.int32))
    ctx = mp.get_context("fork")
    ready, done = ctx.Barrier(procs + 1), ctx.Barrier(procs + 1)
    go = ctx.Value("i", 0)
    counts = ctx.Array("i", procs)
    workers = [ctx.Process(target=_cpu_worker, args=(w, params, args.topology, feats, n_utts_max, frames, ready,
                                                     go, done, counts, threads), daemon=True) for w in range(procs)]
    for w in workers:
        w.start()

    def one_step(n):
        go.value = n
        ready.wait()
        t0 = time.perf_counter()
        done.wait()
        return time.perf_counter() - t0

    t1 = one_step(1)                                                   # calibrate: seconds per utterance per process
    n_per = int(max(1, min(512, round(args.step_seconds / max(t1, 1e-4)))))
    for _ in range(max(args.warmup - 1, 0)):
        one_step(n_per)
    times = [one_step(n_per) for _ in range(args.steps)]
    go.value = 0
    ready.wait()
    for w in workers:
        w.join(timeout=10)
    step_frames = procs * n_per * frames
    total = sum(times)
    value = step_frames * args.steps / total
    sample = ("%d steps x %d processes x %d utterances of %d frames, one utterance per call, %d threads/process"
              % (args.steps, procs, n_per, frames, threads))
    cb = dict(value=round(value, 1), unit=UNIT, cores=procs * threads, kind="port", sample=sample,
              host_cores_visible=cores,
              note="torch-CPU fp32 restatement of the reference forward (TensorFlow 1.x absent: oracle/README in DESIGN.md)")
    out = dict(metric=METRIC, impl="reference", value=round(value, 1), unit=UNIT, n_gpus=args.gpus, steps=args.steps,
               warmup=args.warmup, ms_per_step=round(total / args.steps * 1e3, 3), higher_is_better=True,
               scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
               config=dict(workload="configs[1] sample: %d-frame x %d-dim MFCC utterances, %s, weight set %s"
                                    % (frames, FEAT_DIM, args.topology, args.weight_set),
                           frames_per_step=step_frames),
               cpu_baseline=cb,
               e2e=dict(value=round(value, 1), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--topology", choices=sorted(TOPOLOGIES), default="ModelWithoutDropoutTdnn")
    ap.add_argument("--weight-set", choices=["A", "B"], default="B")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--frames", type=int, default=400)
    ap.add_argument("--option", action="append", default=[], help="xv_set_option name=value (diagnostics)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the configs[4] training-step block")
    ap.add_argument("--no-product", action="store_true", help="skip the ark -> ark product-path block (configs[2] / [3])")
    ap.add_argument("--step-seconds", type=float, default=1.5, help="reference arm: CPU seconds per step (calibrated)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
