#!/usr/bin/env python
"""bench.py — masked frames/s of the quantize-and-predict hot path (VQ assign + masked CE fwd/bwd).

    python bench.py --gpus N --steps K --warmup W            # our arm (B200 kernels)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the CPU restatement (oracle port)

One "step" = one pass of the hot path over one synthetic batch (BASELINE.json configs[1], per GPU):
  VectorQuantizer fwd (assign + gather/straight-through) + EMA codebook update + commitment loss fwd/bwd on
  64 lines x 128 frames x D=256 against an 8192 x 256 codebook, then LinearHead + masked cross-entropy fwd/bwd
  (Dh=512, V=8192 = the codebook's labels, 15 % masking) on the labels the quantizer just produced.
`value`  : whole-job masked frames/s with every input resident in HBM (CUDA-graph replay of the step).
`e2e`    : same metric through the public module API with HOST (pinned) inputs: H2D of features / hidden
           states / mask rows and the D2H loss read are inside the timed region.
`roofline`: the dominant kernel (distance GEMM + arg-min) timed alone with CUDA events.
`cpu_baseline`: the oracle port of the reference path on this box's host cores (rank 0, N=1 only).
Multi-GPU: weak scaling, batch-sharded (every rank owns 64 lines), codebook/head replicated, EMA sums|counts
and head gradients|loss reduced in place inside the step by the library's peer-memory kernel (NVSwitch multimem).

A/B knobs of THIS script (environment variables; the defaults are the measured optimum, profiles/r2_notes.md; the
library itself reads no environment variable in its production build):
  PERO_CE_PRIO (-5) / PERO_EMA_PRIO (-2) / PERO_COMMIT_PRIO (0) / PERO_COMM_PRIO (-5)   stream priorities of the masked-CE
      chain, the EMA chain, the quantize + commitment chain and the gradient-exchange stream
  PERO_STEP_GATHER_SIDE (1), PERO_STEP_COMMIT_SIDE (1), PERO_STEP_PREP_LOW (1), PERO_STEP_SPLIT_EMA (1 on one GPU)
      which chain of the step runs on which stream
  PERO_DP_CHUNKS (1)   label ranges of the data-parallel head backward;  PERO_PEER_BLOCKS (24), PERO_PEER_MULTICAST (auto)
  PERO_E2E_LOSS (pipelined | sync), PERO_E2E_ST (1: single-threaded autograd), PERO_E2E_GRAPH (1), PERO_E2E_BF16 (1)
"""
import argparse
import json
import os

# The step is captured as a CUDA graph with up to nine parallel branches (three compute chains, operand preparation, the
# loss read-out and, data parallel, three exchange chains): with the default of 8 hardware queues two branches can share
# one and serialise.  Set before CUDA is initialised; an explicit setting of the caller wins.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(lines=64, frames=128, K=8192, D=256, Dh=512, V=8192, p=0.15, commitment_cost=0.25, decay=0.99, epsilon=1e-5)
WORKLOAD = ("configs[1]: VQ-VAE quantizer fwd/bwd + EMA update, 8192x256 codebook, 64 lines x 128 frames, "
            "+ masked CE fwd/bwd over the 8192 labels (Dh=512, 15% masking)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="skip the BASELINE configs[2..4] legs (c3, c4, c5)")
    ap.add_argument("--skip-gpu-baseline", action="store_true", help="skip the torch-on-B200 reference leg")
    ap.add_argument("--skip-dp-check", action="store_true", help="N > 1: skip the data-parallel == single-process check")
    ap.add_argument("--cpu-time-cap", type=float, default=150.0, help="reference arm: stop after this many seconds of CPU steps")
    ap.add_argument("--timeline", default=None, help="write the kernel timeline (CUPTI) of 2 step replays to this file")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ data
def make_batch(rank, seed=1236):
    """Seeded synthetic '40-px text-line features' (warmed state): frames are drawn
    around codewords so the index distribution stays spread; hidden states ~ N(0,1); head ~ nn.Linear init."""
    c = CFG
    g = torch.Generator().manual_seed(seed)                 # replicated state: same on every rank
    weight = torch.randn(c["K"], c["D"], generator=g)
    bound = 1.0 / np.sqrt(c["Dh"])
    W = (torch.rand(c["V"], c["Dh"], generator=g) * 2 - 1) * bound
    b = (torch.rand(c["V"], generator=g) * 2 - 1) * bound
    gr = torch.Generator().manual_seed(seed + 1000 + rank)   # per-rank shard of the batch
    N = c["lines"] * c["frames"]
    j = torch.randint(0, c["K"], (N,), generator=gr)
    rows = weight[j] + 0.5 * torch.randn(N, c["D"], generator=gr)
    x = rows.view(c["lines"], c["frames"], c["D"]).permute(0, 2, 1).contiguous().view(c["lines"], c["D"], 1, c["frames"])
    gq = torch.randn(c["lines"], c["D"], 1, c["frames"], generator=gr)
    h = torch.randn(c["lines"], c["frames"], c["Dh"], generator=gr)
    rng = np.random.default_rng(seed + rank)
    mask = (rng.random((c["lines"], c["frames"])) < c["p"]).astype(int)
    return dict(weight=weight, W=W, b=b, x=x, gq=gq, h=h, mask=mask)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 20 ms from before the warm-up replays; samples are then filtered to the
    wall-clock window of the timed region (falling back to warm-up + timed when the region is shorter than
    a few sampling periods)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(rows):
            sm, smax, pw, reasons = [], [], [], set()
            for _, ln in rows:
                f = [t.strip() for t in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return sm, smax, pw, reasons

        inside = [r for r in self.lines if t0 is not None and t0 - 0.02 <= r[0] <= t1 + 0.02]
        window = "timed region"
        if len(inside) < 3:
            inside, window = self.lines, "warm-up + timed region (timed region shorter than 3 sampling periods)"
        sm, smax, pw, reasons = parse(inside)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_step_factory(batch):
    """One step of the same workload restated on the reference's CPU ops (oracle port): VectorQuantizer
    fwd + EMA + commitment loss and gradient, LinearHead over EVERY frame + MaskedCrossEntropyLoss fwd/bwd
    (the reference's head is not mask-aware: masked_pretraining/model.py:60-61, 78-82)."""
    from oracle import pero_oracle as O
    c = CFG
    state = dict(weight=batch["weight"].clone(), ema_w=batch["weight"].clone(), cs=torch.ones(c["K"]))
    mask_t = torch.from_numpy(batch["mask"])

    def step():
        out = O.vq_forward(batch["x"], state["weight"], state["ema_w"], state["cs"], c["decay"], c["epsilon"], True)
        loss_c = O.vq_calculate_loss(out["quantized"], batch["x"], c["commitment_cost"], c["decay"])
        _, g_feat = O.vq_calculate_loss_grads(out["quantized"], batch["x"], c["commitment_cost"], c["decay"])
        g_x = O.vq_forward_grad_inputs(batch["gq"]) + g_feat
        state.update(weight=out["weight"], ema_w=out["ema_w"], cs=out["ema_cluster_size"])
        labels = out["indices"].view(c["lines"], c["frames"])
        loss, d_h, d_W, d_b = O.head_masked_ce(batch["h"], batch["W"], batch["b"], labels, mask_t)
        return float(loss_c) + float(loss), g_x, d_h, d_W, d_b

    return step


def run_cpu(batch, reps, warm, time_cap=None):
    """Returns (mean s/step, min s/step, timed steps done, warm-up steps done).  `time_cap` (seconds) bounds the
    whole call: the warm-up is cut to one step and the timed loop stops early when the cap is reached."""
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_factory(batch)
    t_start = time.perf_counter()
    warm_done = 0
    for _ in range(warm):
        step()
        warm_done += 1
        if time_cap is not None and time.perf_counter() - t_start > 0.25 * time_cap:
            break
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
        if time_cap is not None and time.perf_counter() - t_start > time_cap:
            break
    return float(np.mean(ts)), float(np.min(ts)), len(ts), warm_done


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = make_batch(0)
    M = int(batch["mask"].sum())
    # --steps / --warmup are honoured up to a wall-clock cap (a CPU step takes 0.1-1 s depending on the host)
    mean_s, _, steps, warm = run_cpu(batch, max(1, args.steps), max(0, args.warmup), time_cap=args.cpu_time_cap)
    val = M / mean_s
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": "masked_frames_per_sec", "value": val, "unit": "masked frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": mean_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, **{k: CFG[k] for k in ("lines", "frames", "K", "D", "Dh", "V", "p")},
                       "masked_frames_per_step": float(M), "frames_per_step": CFG["lines"] * CFG["frames"],
                       "parallelism": "host cores of rank 0 (N ranks do not add CPU work)", "exchange": None, "l2": "n/a",
                       "launch": "cpu", "host_affinity": "all host cores"},
            # (kept out of `config`, whose key set is the same as our arm's)
            "reference_arm": {"note": "reference path restated on torch CPU ops (oracle port)", "steps_requested": args.steps,
                              "warmup_requested": args.warmup, "time_cap_s": args.cpu_time_cap},
            "cpu_baseline": {"value": val, "unit": "masked frames/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} full-size steps of one 64-line batch after {warm} warm-up"},
            "e2e": {"value": val, "unit": "masked frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
class DeviceStep:
    """The step on device-resident inputs through the tensor-level wrappers of the C ABI."""

    def __init__(self, batch, dev, dp):
        from pero_pretraining_b200 import ops
        self.ops, self.dp, c = ops, dp, CFG
        self.lines = int(batch["x"].shape[0])          # CFG["lines"], or all ranks' lines in the equivalence check
        self.x = batch["x"].to(dev).view(self.lines, c["D"], c["frames"])
        self.gq = batch["gq"].to(dev).view(self.lines, c["D"], c["frames"])
        self.h = batch["h"].to(dev).view(-1, c["Dh"])
        self.W, self.b = batch["W"].to(dev), batch["b"].to(dev)
        self.weight = batch["weight"].to(dev)
        self.ema_w = self.weight.clone()
        self.cs = torch.ones(c["K"], device=dev)
        rows = np.flatnonzero(batch["mask"].reshape(-1) == 1).astype(np.int32)
        self.M = int(rows.size)
        self.rows = torch.from_numpy(rows).to(dev)
        self.cb = ops.PreparedCodebook(c["K"], c["D"], dev).prepare(self.weight)
        self.packed = torch.empty(self.lines * c["frames"], dtype=torch.int64, device=dev)
        self.head = ops.PreparedHead(c["V"], c["Dh"], dev)
        self.m_global = float(self.M)
        if dp:
            t = torch.tensor([float(self.M)], device=dev)
            torch.distributed.all_reduce(t)
            self.m_global = float(t.item())
        self.out = {}
        # The masked-CE chain is the critical path: its stream (and the exchange stream) outrank the EMA chain, so
        # that pending GEMM CTAs are scheduled ahead of the EMA chain's bandwidth-bound kernels.
        # (data parallel: the EMA chain feeds an exchange that should be out of the way before the gradient exchange
        # needs the links, so there it runs at high priority too)
        ema_prio = int(os.environ.get("PERO_EMA_PRIO", "-2"))
        self.commit_side = os.environ.get("PERO_STEP_COMMIT_SIDE", "1") == "1"
        self.s_commit = torch.cuda.Stream(device=dev, priority=int(os.environ.get("PERO_COMMIT_PRIO", "0")))
        self.s_ema = torch.cuda.Stream(device=dev, priority=ema_prio)
        # the masked-CE chain (and, at the end of the step, its scatter) is the critical path: highest priority the device
        # offers, so that its pending CTAs are placed ahead of the EMA / commitment kernels whenever SM slots free up
        self.s_ce = torch.cuda.Stream(device=dev, priority=int(os.environ.get("PERO_CE_PRIO", "-5")))
        self.s_prep = torch.cuda.Stream(device=dev, priority=0)
        # the gradient exchange is the data-parallel tail: its CTAs must not queue behind pending GEMM CTAs of the CE chain
        self.s_comm = torch.cuda.Stream(device=dev, priority=int(os.environ.get("PERO_COMM_PRIO", "-5")))
        # Data-parallel exchange ranges (EMA sums|counts and d_W|d_b|loss_sum) live in a peer-mapped buffer and
        # are reduced in place by the library's own NVLink/NVSwitch kernel.
        self.peer = self.ema_x = self.grad_x = None
        self.n_ema = c["K"] * c["D"] + c["K"]
        self.n_grad = c["V"] * c["Dh"] + c["V"]
        if dp:
            from pero_pretraining_b200.peer import PeerBuffer, PeerRange
            blocks = int(os.environ.get("PERO_PEER_BLOCKS", "24"))
            mc = os.environ.get("PERO_PEER_MULTICAST", "auto")
            mc = None if mc == "auto" else mc != "0"
            # one buffer (= one set of barrier words) per exchange chain: the EMA exchange and the gradient exchange
            # run on their own streams and may overlap
            self.peer_ema = PeerBuffer(4 * self.n_ema + 1024, dev, n_blocks=blocks, use_multicast=mc)
            self.peer = PeerBuffer(4 * self.n_grad + 1024, dev, n_blocks=blocks, use_multicast=mc)
            self.ema_x = PeerRange(self.peer_ema, self.n_ema, torch.float32)
            self.grad_x = PeerRange(self.peer, c["V"] * c["Dh"], torch.float32)             # d_W
            # d_b | loss_sum travel as a small message of their own (own buffer = own barrier words: it may overlap the
            # d_W exchange), so that the 16.8 MB exchange starts the moment the d_W GEMM is done
            self.peer_db = PeerBuffer(4 * (c["V"] + 4) + 1024, dev, n_blocks=2, use_multicast=mc)
            self.db_x = PeerRange(self.peer_db, c["V"] + 4, torch.float32)
            self.s_comm_ema = torch.cuda.Stream(device=dev, priority=-1)
            self.s_comm_db = torch.cuda.Stream(device=dev, priority=-1)
            # label-axis ranges of the head backward: each range's d_W rows are exchanged while the next is computed
            chunks = max(1, int(os.environ.get("PERO_DP_CHUNKS", "1")))
            step = max(256, (c["V"] // chunks + 255) // 256 * 256)
            self.v_ranges = [(v, min(v + step, c["V"])) for v in range(0, c["V"], step)]

    def __call__(self):
        """Three chains, forked onto side streams (captured as parallel branches of the CUDA graph):
          A (capture stream)  frame preparation + distance GEMM, unpack, gather/straight-through, commitment loss fwd/bwd
          B (EMA stream)      EMA codebook update (needs A's gather to have read the old codebook before it is overwritten)
          C (CE stream)       head operand preparation and the gather of the masked hidden states at the START of the step
                              (neither needs this step's labels), then -- directly behind the distance GEMM, reading the
                              labels from its packed (distance, index) winners -- logits GEMM + LSE, dlogits, d_W | d_h,
                              scatter; the loss sum is read out of the partials at the end, off the GEMM chain."""
        ops, c = self.ops, CFG
        N = self.lines * c["frames"]
        main = torch.cuda.current_stream()
        s_ema, s_ce = self.s_ema, self.s_ce
        s_ema.wait_stream(main)
        # The frame preparation in front of the distance GEMM is the head of the critical path: it is issued first, at
        # high priority; the head's operand preparation and the gather of the masked hidden states (needed only behind
        # the distance GEMM) go to a low-priority stream and fill in beside it.
        s_prep = self.s_prep if os.environ.get("PERO_STEP_PREP_LOW", "1") == "1" else s_ce
        s_prep.wait_stream(main)
        packed = self.packed                         # reset by the frame preparation pass of vq_assign
        with torch.cuda.stream(s_prep):
            ce_ws = ops.masked_ce_gather(self.h, self.rows, c["V"])
            self.head.prepare(self.W, self.b)        # head weights change every optimizer step in training
        _, _, x_rows = ops.vq_assign(self.x, self.cb, self.lines, c["frames"], True, want_rows=True, packed=packed,
                                     init_packed=True)
        s_ce.wait_stream(s_prep)
        assigned = torch.cuda.Event()
        assigned.record(main)
        s_ce.wait_event(assigned)
        # --- chain A (main stream): unpack, EMA sums (their exchange goes to the communication stream at once), quantize +
        # commitment loss fwd/bwd, EMA apply.  Data parallel: ONE stream, EMA sums first, so that their exchange is out of
        # the way before the gradient exchange needs the links (175-178 us per step at 2 GPUs; 186 with the EMA update on
        # its own stream, 189 with EMA sums first and the commitment chain on a side stream: measured, round 2).  Single
        # GPU: the EMA update runs on its own low-priority stream beside the rest (133 instead of 143 us).
        idx, _ = ops.vq_unpack(packed)
        split_ema = os.environ.get("PERO_STEP_SPLIT_EMA", "0" if self.dp else "1") == "1"
        if split_ema:
            s_ema.wait_stream(main)
        with torch.cuda.stream(s_ema if split_ema else main):
            if not self.dp:
                sums = ops.vq_ema_accumulate(x_rows, idx, c["K"])
            else:
                sums = ops.vq_ema_accumulate(x_rows, idx, c["K"], out=self.ema_x.tensor)
                self.s_comm_ema.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.s_comm_ema):
                    self.ema_x.all_reduce_sum_()
        # quantize + commitment loss fwd/bwd: a short chain of bandwidth-bound kernels nothing else in the step waits for
        # (except the EMA apply, which must not overwrite the codebook before the gather has read it).  Single GPU: on a
        # low-priority stream of its own, so that the longer EMA chain gets the SM slots beside the GEMMs first.
        # quantized output and the commitment loss value in one pass (the loss needs mean((q - x)^2) only): on the main
        # stream, right behind the distance GEMM -- the EMA apply may not overwrite the codebook before this has read it
        s_q = self.s_commit if (split_ema and self.commit_side) else main
        gather_side = s_q is not main and os.environ.get("PERO_STEP_GATHER_SIDE", "1") == "1"
        if gather_side:
            s_q.wait_stream(main)
        with torch.cuda.stream(s_q if gather_side else main):
            q, loss_c = ops.vq_gather_st_mse(x_rows, idx, self.weight, self.lines, c["frames"], True, 0.0, c["commitment_cost"])
            gathered = torch.cuda.Event()
            gathered.record(torch.cuda.current_stream())
        if s_q is not main and not gather_side:
            s_q.wait_stream(main)
        with torch.cuda.stream(s_q):
            g_x = ops.vq_st_commit_bwd(self.gq, q, self.x, 2.0 * c["commitment_cost"] / q.numel())
        with torch.cuda.stream(s_ema if split_ema else main):
            if self.dp:
                torch.cuda.current_stream().wait_stream(self.s_comm_ema)
            if split_ema:
                s_ema.wait_event(gathered)
            ops.vq_ema_apply(sums, self.ema_w, self.cs, self.weight, c["decay"], c["epsilon"], self.cb)
        # --- chain C
        Dh = c["Dh"]
        with torch.cuda.stream(s_ce):
            _, _, ws = ops.masked_ce_fwd(self.h, self.rows, packed, self.head, ws=ce_ws, finalize=False, labels_packed=True,
                                         keep_logits=True)
            if not self.dp:
                # the loss is read out of the log-sum-exp partials beside the backward GEMMs, not behind them
                self.s_comm.wait_stream(s_ce)
                with torch.cuda.stream(self.s_comm):
                    loss_sum, lse = ops.masked_ce_loss(ws, N, Dh, self.M, c["V"])
                d_h, d_W, d_b, flat = ops.masked_ce_bwd(self.h, self.rows, packed, self.head, None, None, 1.0 / self.m_global, ws=ws,
                                                        return_flat=True, ws_from_fwd=True, labels_packed=True, logits_in_ws=True)
            else:
                # loss_sum rides in the same exchange range as d_W | d_b.  The backward walks the label axis range by
                # range: each range's rows of d_W are reduced over the ranks on the communication stream while the
                # next range (and finally d_h) is computed.
                g, gb = self.grad_x.tensor, self.db_x.tensor
                V = c["V"]
                # the loss sum is read out of the log-sum-exp partials on a communication stream, beside the backward
                # GEMMs, straight into its slot of the small exchange range (d_b | loss_sum)
                self.s_comm_db.wait_stream(s_ce)
                with torch.cuda.stream(self.s_comm_db):
                    loss_sum, lse = ops.masked_ce_loss(ws, N, Dh, self.M, V, loss_out=gb[V:V + 1])
                # phase 1 per label range: dlogits + d_W; each range's rows of d_W are reduced over the ranks on the
                # communication stream while the next range (and finally d_h | d_b) is computed
                for v0, v1 in self.v_ranges:
                    _, d_W, _ = ops.masked_ce_bwd(self.h, self.rows, packed, self.head, None, None, 1.0 / self.m_global,
                                                  ws=ws, want_dh=False, want_db=False, dw_out=g.view(V, Dh),
                                                     ws_from_fwd=True, v_range=(v0, v1), labels_packed=True, logits_in_ws=True)
                    self.s_comm.wait_stream(s_ce)
                    with torch.cuda.stream(self.s_comm):
                        self.peer.all_reduce_sum_(self.grad_x.offset + 4 * v0 * Dh, (v1 - v0) * Dh)
                # phase 2: d_h and d_b from the dlogits in the workspace, then the small exchange
                d_h, _, d_b = ops.masked_ce_bwd(self.h, self.rows, packed, self.head, None, None, 1.0 / self.m_global, ws=ws,
                                                want_dw=False, want_db=True, db_out=gb[:V], ws_from_fwd=True, labels_packed=True,
                                                logits_in_ws=True)
                self.s_comm_db.wait_stream(s_ce)
                with torch.cuda.stream(self.s_comm_db):
                    self.db_x.all_reduce_sum_()
                flat = g
        main.wait_stream(s_ema)
        main.wait_stream(s_ce)
        main.wait_stream(self.s_comm)
        if s_q is not main:
            main.wait_stream(s_q)
        if self.dp:
            main.wait_stream(self.s_comm_ema)
            main.wait_stream(self.s_comm_db)
        self.out = dict(idx=idx, x_rows=x_rows, q=q, loss_c=loss_c, g_x=g_x, sums=sums, loss_sum=loss_sum, lse=lse, ws=ws,
                        d_h=d_h, d_W=d_W, d_b=d_b, flat=flat)
        return self.out


def _compare_with_single_process(dev, world, mine, all_idx, replicas_identical):
    """Rank 0: the single-GPU step on all ranks' lines (regenerated from their seeds) against the data-parallel results."""
    c = CFG
    parts = [make_batch(r) for r in range(world)]
    cat = dict(weight=parts[0]["weight"], W=parts[0]["W"], b=parts[0]["b"],
               x=torch.cat([p["x"] for p in parts]), gq=torch.cat([p["gq"] for p in parts]),
               h=torch.cat([p["h"] for p in parts]), mask=np.concatenate([p["mask"] for p in parts], axis=0))
    single = DeviceStep(cat, dev, False)
    ref = single()
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
    res = {"world": world, "frames": int(single.lines * c["frames"]), "masked_frames": int(single.M),
           "replicas_bit_identical": replicas_identical,
           "indices_equal": bool(torch.equal(torch.cat(all_idx), ref["idx"])),
           "codebook_after_ema_max_rel_err": rel(mine["weight"], single.weight),
           "d_W_max_err_over_max": rel(mine["d_W"], ref["d_W"].reshape(-1)),
           "d_b_max_err_over_max": rel(mine["d_b"], ref["d_b"]),
           "loss_sum_rel_err": rel(mine["loss"], ref["loss_sum"].reshape(1))}
    res["ok"] = bool(res["replicas_bit_identical"] and res["indices_equal"] and res["codebook_after_ema_max_rel_err"] < 1e-4
                     and res["d_W_max_err_over_max"] < 2e-2 and res["d_b_max_err_over_max"] < 2e-2
                     and res["loss_sum_rel_err"] < 1e-3)
    return res


def dp_equivalence_check(dev, rank, world):
    """Data-parallel step == single-process step on the concatenated batch, checked inside the multi-GPU bench
    run so that the driver's own N > 1 runs carry the evidence (the -m gpu suite's multi-GPU test is skipped on a 1-GPU
    box).  Every rank runs ONE data-parallel step from the bench's initial state; rank 0 also runs the single-GPU step on
    all ranks' lines (regenerated from their seeds) and compares: indices exactly, the EMA-updated codebook, d_W, d_b and
    the loss sum within fp32 re-association / bf16-operand tolerance; replicas must hold identical bits."""
    try:
        c = CFG
        ds = DeviceStep(make_batch(rank), dev, True)
        out = ds()
        torch.cuda.synchronize()
        mine = dict(idx=out["idx"], weight=ds.weight, d_W=out["d_W"].reshape(-1), d_b=out["d_b"], loss=out["loss_sum"].reshape(1))
        # replicas: the exchanged quantities and the updated codebook carry the same bits on every rank
        sums = torch.stack([mine[k].contiguous().view(torch.int32).to(torch.int64).sum() for k in ("weight", "d_W", "d_b", "loss")])
        gathered = [torch.empty_like(sums) for _ in range(world)]
        torch.distributed.all_gather(gathered, sums)
        replicas_identical = all(bool(torch.equal(g, gathered[0])) for g in gathered)
        all_idx = [torch.empty_like(mine["idx"]) for _ in range(world)]
        torch.distributed.all_gather(all_idx, mine["idx"])
        res = None
        if rank == 0:
            try:      # a failure of the rank-0 comparison must not keep rank 0 away from the barrier below
                res = _compare_with_single_process(dev, world, mine, all_idx, replicas_identical)
            except Exception as e:      # noqa: BLE001
                res = {"error": f"{type(e).__name__}: {e}"}
        del ds, out, mine
        torch.cuda.synchronize()
        torch.distributed.barrier()
        torch.cuda.empty_cache()
        return res
    except Exception as e:      # noqa: BLE001  (the check must never take the measurement down)
        return {"error": f"{type(e).__name__}: {e}"}


def count_kernels(fn, align=None):
    """Kernels launched by one step (CUPTI via torch.profiler); -1 when the profiler is unavailable.
    `align` (data parallel: a barrier) runs after the profiler has started, so that ranks whose profiler start-up
    differs by seconds enter the step's exchange kernels together."""
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            if align is not None:
                align()
                torch.cuda.synchronize()
            fn()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower()
                 and "nccl" not in e.name.lower()]
        ours = [n for n in names if "pero" in n or "cub" in n.lower() or "allreduce" in n]
        return len(names), len(ours)
    except Exception:
        return -1, -1


def dump_timeline(run, path):
    """Kernel start/duration/stream of 2 replays of the step (diagnostics only; not a timed number)."""
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if "cuda" in str(getattr(e, "device_type", "")).lower()]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start if evs else 0
    with open(path, "w") as f:
        for e in evs:
            f.write(f"{e.time_range.start - t0:10.1f} {e.time_range.end - e.time_range.start:8.1f} {e.name[:110]}\n")


def e2e_leg(batch, dev, dp, steps, warm, repeats=5, loss_read="pipelined", single_thread_autograd=True):
    """Public module API, host inputs: per step H2D of x (fp32) / h (bf16: what --bfloat16 training hands the head,
    masked_pretraining/trainer.py:57-59) / mask rows from pinned memory, D2H of the loss.  `repeats` timed loops of
    `steps` steps each; returns (median s/step, min s/step, all, h2d bytes, d2h bytes, last loss).
    loss_read: "sync" = `loss.item()` at the end of every step (the host waits for the device before it starts launching
    the next step); "pipelined" = every step's loss is copied D2H into pinned memory asynchronously and read by the host
    one step later (the usual logging pattern of a training loop: the host keeps launching while the device finishes);
    every step's loss is still read inside the timed region, the last one behind the closing synchronize.
    single_thread_autograd: `torch.autograd.set_multithreading_enabled(False)` -- the backward of this step is three
    custom nodes on one device; running them in the calling thread saves the hand-off to the engine's device thread."""
    from pero_pretraining_b200 import LinearHead, VectorQuantizer
    c = CFG
    vq = VectorQuantizer(c["K"], c["D"], c["commitment_cost"], c["decay"], c["epsilon"]).to(dev).train()
    head = LinearHead(c["Dh"], c["V"]).to(dev)
    with torch.no_grad():
        vq.embedding.weight.copy_(batch["weight"]); vq.ema_w.copy_(batch["weight"]); vq.ema_cluster_size.fill_(1.0)
        head.linear.weight.copy_(batch["W"]); head.linear.bias.copy_(batch["b"])
    group = torch.distributed.group.WORLD if dp else None
    if dp:
        vq.enable_data_parallel(group)          # peer-memory exchange of the EMA sums|counts
        head.enable_peer_exchange(group, alias_grads=True)      # ... and of d_W | d_b (every backward is followed by a reset)
    if os.environ.get("PERO_E2E_GRAPH", "1") != "0":
        vq.enable_cuda_graph()                  # the forward's kernel sequence as one graph replay (public opt-in)
    h_src = batch["h"].bfloat16() if os.environ.get("PERO_E2E_BF16", "1") != "0" else batch["h"]
    x_host, h_host = batch["x"].pin_memory(), h_src.pin_memory()
    gq = batch["gq"].to(dev)
    mask = batch["mask"]
    n_rows = int((mask.reshape(-1) == 1).sum())
    h2d = x_host.numel() * 4 + h_host.numel() * h_host.element_size() + n_rows * 4
    loss_val = None
    # Double-buffered input staging: the H2D copy of step i+1 runs on a copy stream while step i computes;
    # every step still waits for ITS OWN inputs to arrive and reads ITS OWN loss back.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(x_host, device=dev), torch.empty_like(h_host, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def stage(i):
        xb, hb = bufs[i & 1]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])            # the step that last used this buffer pair is done
            xb.copy_(x_host, non_blocking=True)
            hb.copy_(h_host, non_blocking=True)
            ready[i & 1].record(copy_stream)

    state = {"i": 0}
    for e in consumed:
        e.record()
    stage(0)
    labels_shape = (c["lines"], c["frames"])
    pipelined = loss_read == "pipelined"
    one = torch.ones((), dtype=torch.float32, device=dev)
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]

    def step():
        i = state["i"]
        state["i"] = i + 1
        main = torch.cuda.current_stream()
        main.wait_event(ready[i & 1])
        stage(i + 1)                                           # prefetch the next step's inputs
        x = bufs[i & 1][0].detach().requires_grad_(True)
        h = bufs[i & 1][1].detach().requires_grad_(True)
        q, idx = vq(x)
        loss_q = vq.calculate_loss(q, x)
        loss_h = head.masked_loss(h, idx.view(labels_shape), mask, None, group)
        head.linear.weight.grad = None
        head.linear.bias.grad = None
        # d(loss_q + loss_h): the two losses are differentiated as two roots with a preallocated unit gradient (no
        # AddBackward node, no ones_like per step); their sum, the step's result, is formed on the detached values
        torch.autograd.backward([loss_q, loss_h, q], [one, one, gq])
        loss = loss_q.detach() + loss_h.detach()
        consumed[i & 1].record(main)
        if not pipelined:
            return float(loss.item())                  # D2H read of the step's result
        loss_host[i & 1].copy_(loss.detach(), non_blocking=True)      # D2H of this step's result, read one step later
        loss_ready[i & 1].record(main)
        if i == 0:
            return None
        loss_ready[(i - 1) & 1].synchronize()
        return float(loss_host[(i - 1) & 1])

    def last_loss():
        i = state["i"] - 1
        loss_ready[i & 1].synchronize()
        return float(loss_host[i & 1])

    prev_mt = torch.autograd.is_multithreading_enabled()
    if single_thread_autograd:
        torch.autograd.set_multithreading_enabled(False)
    for _ in range(max(warm, 10)):            # allocator pools, pinned staging ring and graph caches settle in the first steps
        step()
    times = []
    for _ in range(repeats):
        torch.cuda.synchronize()
        if dp:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            loss_val = step()
        if pipelined:
            loss_val = last_loss()                     # the last step's loss: read inside the timed region too
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dp:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            dt = float(t.item())
        times.append(dt / steps)
    torch.autograd.set_multithreading_enabled(prev_mt)
    return float(np.median(times)), float(min(times)), times, h2d, 4, loss_val


def gemm_roofline_leg(ds, dev, iters, flush):
    """The dominant kernel alone: distance GEMM + arg-min on prepared bf16 frames.  The launches rotate through
    R independent (frames, codebook) sets whose total size exceeds the 126 MB L2 (inputs larger than L2: every
    launch reads its operands from HBM), R launches back to back between two CUDA events on the launching stream;
    per-launch duration = elapsed / R, so event and launch latency are not billed to the kernel."""
    from pero_pretraining_b200 import _lib, ops
    c = CFG
    L = _lib.lib()
    N = c["lines"] * c["frames"]
    per_set = N * c["D"] * 2 + ds.cb.nbytes + N * 8
    R = int(np.ceil(160e6 / per_set))                     # > 126 MB of distinct operands in flight
    g = torch.Generator(device=dev).manual_seed(7)
    base = ds.x.permute(0, 2, 1).reshape(N, c["D"]).contiguous()
    xbs, cbs, packs = [], [], []
    for r in range(R):
        xbs.append((base + 0.01 * torch.randn(N, c["D"], device=dev, generator=g)).bfloat16())
        w = ds.weight + 0.01 * torch.randn(c["K"], c["D"], device=dev, generator=g)
        cbs.append(ops.PreparedCodebook(c["K"], c["D"], dev).prepare(w))
        packs.append(torch.empty(N, dtype=torch.int64, device=dev))
    stream = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(iters + 3):
        for r in range(R):
            L.pero_vq_packed_init(packs[r].data_ptr(), N, stream)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(R):
            rc = L.pero_vq_assign_bf16(xbs[r].data_ptr(), N, c["K"], c["D"], cbs[r].blob.data_ptr(), 0, packs[r].data_ptr(), stream)
            _lib.check(rc, "pero_vq_assign_bf16")
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) / R)
    return float(np.mean(ts)), float(np.min(ts)), R


# ------------------------------------------------------------------------------------------------ GPU bar + other configs
def timed_loop(fn, steps, warm, flush, dp, use_graph=True):
    """ms per call of fn(): `warm` untimed calls, then `steps` calls each bracketed by CUDA events on the current stream
    with the L2 flushed in between; max over ranks.  fn is captured into a CUDA graph when possible."""
    run, graph = fn, None
    for _ in range(max(1, warm)):
        fn()
    torch.cuda.synchronize()
    if use_graph:
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                fn()
            graph.replay()
            torch.cuda.synchronize()
            run = graph.replay
        except Exception as e:      # noqa: BLE001
            sys.stderr.write(f"[bench] graph capture unavailable for {getattr(fn, '__name__', fn)}: {e}\n")
            torch.cuda.synchronize()
            run, graph = fn, None
    for _ in range(max(3, warm)):
        run()
    torch.cuda.synchronize()
    if dp:
        torch.distributed.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in ev:
        flush.zero_()
        e0.record()
        run()
        e1.record()
    torch.cuda.synchronize()
    if dp:
        torch.distributed.barrier()
    ts = [e0.elapsed_time(e1) for e0, e1 in ev]
    total = float(sum(ts))
    if dp:
        t = torch.tensor([total], device=flush.device, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        total = float(t.item())
    return total / steps, float(min(ts)), ("cuda_graph" if graph is not None else "eager"), graph


def gpu_baseline_leg(batch, dev, flush, steps=10, warm=3):
    """The bar SURVEY 2b / 8d names: the reference's own torch op sequence for this step run on the SAME B200 (cuBLAS /
    ATen pick their sm_100 kernels), restated here op for op:
      VectorQuantizer.forward + calculate_loss   models/autoencoders.py:193-241  (3 dense fp32 GEMMs, one-hot matrix)
      LinearHead over EVERY frame                masked_pretraining/model.py:104-105
      MaskedCrossEntropyLoss                     masked_pretraining/model.py:78-82  (boolean-mask gather, F.cross_entropy)
      loss.backward()                            autograd
    fp32 with TF32 off (torch's default, what the reference runs), and with the head + loss under bf16 autocast as
    masked_pretraining/trainer.py:57-59 does with --bfloat16.  Device time per step by CUDA events, L2 flushed."""
    import torch.nn.functional as F
    c = CFG
    K = c["K"]
    x = batch["x"].to(dev).requires_grad_(True)                       # [Nl, D, 1, T]
    gq = batch["gq"].to(dev)
    h = batch["h"].to(dev).requires_grad_(True)                       # [Nl, T, Dh]
    W = batch["W"].to(dev).requires_grad_(True)
    b = batch["b"].to(dev).requires_grad_(True)
    mask_t = torch.from_numpy(batch["mask"]).to(dev)
    state = dict(weight=batch["weight"].to(dev).clone(), ema_w=batch["weight"].to(dev).clone(), cs=torch.ones(K, device=dev))
    out = {}

    def step(autocast):
        weight = state["weight"]
        inputs = x.permute(0, 2, 3, 1).contiguous()                                                    # :205
        flat = inputs.view(-1, c["D"])                                                                 # :209
        distances = (torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(weight ** 2, dim=1)
                     - 2 * torch.matmul(flat, weight.t()))                                             # :212-214
        idx = torch.argmin(distances, dim=1).unsqueeze(1)                                              # :217
        enc = torch.zeros(idx.shape[0], K, device=dev)
        enc.scatter_(1, idx, 1)                                                                        # :218-219
        quantized = torch.matmul(enc, weight).view(inputs.shape)                                       # :222
        with torch.no_grad():                                                                          # :225-237
            cs = state["cs"] * c["decay"] + (1 - c["decay"]) * torch.sum(enc, 0)
            n = torch.sum(cs)
            cs = (cs + c["epsilon"]) / (n + K * c["epsilon"]) * n
            dw = torch.matmul(enc.t(), flat)
            ema_w = state["ema_w"] * c["decay"] + (1 - c["decay"]) * dw
            state.update(weight=ema_w / cs.unsqueeze(1), ema_w=ema_w, cs=cs)
        quantized = inputs + (quantized - inputs).detach()                                             # :239
        q = quantized.permute(0, 3, 1, 2).contiguous()                                                 # :241
        loss_c = c["commitment_cost"] * F.mse_loss(q.detach(), x)                                      # :198-200
        labels = idx.detach().view(c["lines"], c["frames"])
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):                           # trainer.py:57-59
            logits = F.linear(h, W, b)                                                                 # model.py:104-105
            sel = mask_t == 1                                                                          # model.py:79-80
            loss = F.cross_entropy(logits[sel], labels[sel])                                           # model.py:82
        x.grad = h.grad = W.grad = b.grad = None
        torch.autograd.backward([loss_c + loss, q], [None, gq])
        out["loss"] = loss

    res = {}
    M = int(batch["mask"].sum())
    for name, autocast in (("fp32", False), ("bf16_autocast_head", True)):
        state.update(weight=batch["weight"].to(dev).clone(), ema_w=batch["weight"].to(dev).clone(), cs=torch.ones(K, device=dev))
        ms, ms_min, _, _ = timed_loop(lambda: step(autocast), steps, warm, flush, False, use_graph=False)
        res[name] = {"ms_per_step": ms, "ms_min": ms_min, "value": M / (ms * 1e-3), "unit": "masked frames/s"}
    res["kind"] = "port"
    res["what"] = ("the reference's torch op sequence (models/autoencoders.py:193-241, masked_pretraining/model.py:78-105, "
                   "autograd backward) restated in bench.py and run on this B200 through cuBLAS/ATen; TF32 off; eager launches")
    res["steps"] = steps
    del x, h, W, b, state
    torch.cuda.empty_cache()
    return res


class CeChain:
    """Head operand preparation + fused masked CE forward + backward (+ the data-parallel gradient exchange) on
    device-resident inputs: BASELINE configs[2] (V = 4096 labels, Dh = 512, 128 frames per line, 15 % masking)."""

    def __init__(self, dev, dp, Nl, T_, Dh, V, p, rank, seed, h_dtype):
        from pero_pretraining_b200 import ops
        self.ops, self.dp, self.V, self.Dh = ops, dp, V, Dh
        g = torch.Generator().manual_seed(seed)
        bound = 1.0 / np.sqrt(Dh)
        self.W = ((torch.rand(V, Dh, generator=g) * 2 - 1) * bound).to(dev)
        self.b = ((torch.rand(V, generator=g) * 2 - 1) * bound).to(dev)
        gr = torch.Generator().manual_seed(seed + 1000 + rank)
        self.h = torch.randn(Nl * T_, Dh, generator=gr).to(dev).to(h_dtype)
        self.labels = torch.randint(0, V, (Nl * T_,), generator=gr).to(dev)
        mask = (np.random.default_rng(seed + rank).random((Nl, T_)) < p)
        rows = np.flatnonzero(mask.reshape(-1)).astype(np.int32)
        self.M = int(rows.size)
        self.rows = torch.from_numpy(rows).to(dev)
        self.head = ops.PreparedHead(V, Dh, dev)
        self.m_global = float(self.M)
        self.n_grad = V * Dh + V
        if dp:
            from pero_pretraining_b200.peer import PeerBuffer, PeerRange
            t = torch.tensor([float(self.M)], device=dev)
            torch.distributed.all_reduce(t)
            self.m_global = float(t.item())
            self.peer = PeerBuffer(4 * self.n_grad + 1024, dev)
            self.grad_x = PeerRange(self.peer, self.n_grad + 1, torch.float32)

    def __call__(self):
        ops = self.ops
        self.head.prepare(self.W, self.b)
        if not self.dp:
            loss_sum, lse, ws = ops.masked_ce_fwd(self.h, self.rows, self.labels, self.head, keep_logits=True)
            self.out = ops.masked_ce_bwd(self.h, self.rows, self.labels, self.head, lse, None, 1.0 / self.m_global, ws=ws,
                                         ws_from_fwd=True, logits_in_ws=True)
        else:
            g = self.grad_x.tensor
            loss_sum, lse, ws = ops.masked_ce_fwd(self.h, self.rows, self.labels, self.head, loss_out=g[self.n_grad:],
                                                  keep_logits=True)
            self.out = ops.masked_ce_bwd(self.h, self.rows, self.labels, self.head, lse, None, 1.0 / self.m_global, ws=ws,
                                         ws_from_fwd=True, flat_out=g[:self.n_grad], return_flat=True, logits_in_ws=True)
            self.grad_x.all_reduce_sum_()
        return self.out


def vqvae_quantize_leg(dev, flush, steps, lines=64, frames=128, C=256, K=8192, D=256):
    import copy
    from pero_pretraining_b200 import VQVAE

    class _Pass(torch.nn.Module):
        def __init__(self, c):
            super().__init__()
            self.out_channels = self.base_channels = c

        def forward(self, x):
            return x

    torch.manual_seed(1240)
    fused = VQVAE(_Pass(C), _Pass(C), K, D, 0.25, 0.99).to(dev).eval()
    plain = copy.deepcopy(fused)
    fused.fuse_projections, plain.fuse_projections = True, False
    feats = torch.randn(lines, C, 1, frames, device=dev)
    res = {}
    for name, m in (("fused", fused), ("conv2d", plain)):
        def step(m=m):
            with torch.no_grad():
                return m.quantize(feats)
        ms, ms_min, launch, graph = timed_loop(step, steps, 3, flush, False)
        res[name] = {"ms_per_step": ms, "ms_min": ms_min, "launch": launch, "frames_per_s": lines * frames / (ms * 1e-3)}
        graph = None
    with torch.no_grad():
        la, lb = fused.quantize(feats)[1], plain.quantize(feats)[1]
    N = lines * frames
    fl = 2.0 * N * (C * D + K * D)                                 # encoder projection + assign
    return {"workload": f"VQVAE.quantize, eval (label production: the decoder-projected codebook is computed once per codebook version): {lines} lines x {frames} frames, {C}->{D} projection, {K}x{D} codebook, {D}->{C} projection",
            "fused": res["fused"], "conv2d_projections": res["conv2d"],
            "speedup": res["conv2d"]["ms_per_step"] / res["fused"]["ms_per_step"],
            "labels_equal_fraction": float((la == lb).float().mean()),
            "fused_tflops_algorithmic": fl / (res["fused"]["ms_per_step"] * 1e-3) / 1e12}


def configs_legs(dev, dp, rank, world, flush, steps):
    """BASELINE configs[2..4] measured with the same timing method as the headline workload:
      c3  masked CE over 4096 labels, Dh 512, 32 lines x 128 frames PER GPU, bf16 hidden states, data parallel (weak)
      c4  assign + quantize + EMA update, 16384 x 512 codebook, 512 lines x 128 frames in total, batch-sharded (strong)
      c5  codebook-sharded assign, 65536 x 512 codebook split over the ranks, 2**20 frames on every rank, MIN exchange
    Every entry: ms per step (device events, L2 flushed, max over ranks), throughput, algorithmic TFLOP/s and the
    fraction of the measured bf16 peak (burst for the single-kernel c5, sustained for the multi-kernel steps)."""
    from pero_pretraining_b200 import ShardedCodebook, VectorQuantizer
    burst, sustained, hbm, src = peaks()
    out = {}
    # ---- c3
    ce = CeChain(dev, dp, 32, 128, 512, 4096, 0.15, rank, 1237, torch.bfloat16)
    ms, ms_min, launch, graph = timed_loop(ce, steps, 3, flush, dp)
    fl = 6.0 * ce.m_global * 512 * 4096
    out["c3"] = {"workload": "configs[2]: masked CE fwd+bwd, V=4096, Dh=512, 32 lines x 128 frames per GPU, 15% masking, bf16 hidden states",
                 "n_gpus": world, "scaling": "weak", "ms_per_step": ms, "ms_min": ms_min, "masked_frames_per_step": ce.m_global,
                 "value": ce.m_global / (ms * 1e-3), "unit": "masked frames/s", "tflops": fl / (ms * 1e-3) / 1e12,
                 "frac_of_sustained": fl / (ms * 1e-3) / 1e12 / sustained / world, "launch": launch,
                 "note": "4-6 GFLOP per GPU: launch/latency-bound (SURVEY 7), reported as measured"}
    graph = None
    del ce
    # ---- c4
    K, D, lines, T_ = 16384, 512, 512, 128
    my_lines = lines // world
    g = torch.Generator().manual_seed(1238)
    w0 = torch.randn(K, D, generator=g)
    gd = torch.Generator(device=dev).manual_seed(1238 + rank)
    j = torch.randint(0, K, (my_lines * T_,), device=dev, generator=gd)
    w0d = w0.to(dev)
    rows = w0d[j] + 0.5 * torch.randn(my_lines * T_, D, device=dev, generator=gd)
    x = rows.view(my_lines, 1, T_, D).permute(0, 3, 1, 2).contiguous()
    del rows
    vq = VectorQuantizer(K, D, 0.25, 0.99).to(dev).train()
    with torch.no_grad():
        vq.embedding.weight.copy_(w0d); vq.ema_w.copy_(w0d); vq.ema_cluster_size.fill_(1.0)
    if dp:
        vq.enable_data_parallel()

    def c4_step():
        with torch.no_grad():
            return vq(x)

    ms, ms_min, launch, graph = timed_loop(c4_step, steps, 3, flush, dp)
    fl = 2.0 * lines * T_ * K * D
    out["c4"] = {"workload": "configs[3]: PQ-AE feature quantization: assign + quantize + EMA codebook update, 16384x512 codebook, 512 lines x 128 frames in total",
                 "n_gpus": world, "scaling": "strong", "ms_per_step": ms, "ms_min": ms_min, "frames_per_step": lines * T_,
                 "value": lines * T_ / (ms * 1e-3), "unit": "frames/s", "tflops": fl / (ms * 1e-3) / 1e12,
                 "frac_of_sustained": fl / (ms * 1e-3) / 1e12 / sustained / world, "launch": launch,
                 "api": "VectorQuantizer.forward (training mode)" + (" + enable_data_parallel()" if dp else "")}
    graph = None
    del vq, x, w0d
    torch.cuda.empty_cache()
    # ---- c5
    K, D, N = 65536, 512, 1 << 20
    gd = torch.Generator(device=dev).manual_seed(1239)            # every rank holds ALL frames: same seed everywhere
    C = torch.randn(K, D, device=dev, generator=gd)
    X = torch.randn(N, D, device=dev, generator=gd)
    sc = ShardedCodebook(C, K, rank, world, peer_frames=N if dp else 0)
    del C

    def c5_step():
        return sc.assign(X, N, 1, False)

    ms, ms_min, launch, graph = timed_loop(c5_step, max(3, steps // 4), 2, flush, dp)
    fl = 2.0 * N * K * D
    out["c5"] = {"workload": "configs[4]: codebook-sharded stress, 65536x512 codebook, 2**20 frames, (distance,index) MIN exchange",
                 "n_gpus": world, "scaling": "strong (codebook split over the ranks)", "ms_per_step": ms, "ms_min": ms_min,
                 "frames_per_step": N, "value": N / (ms * 1e-3), "unit": "frames/s", "tflops": fl / (ms * 1e-3) / 1e12,
                 "frac_of_burst": fl / (ms * 1e-3) / 1e12 / burst / world, "launch": launch,
                 "exchange": (f"pero_peer_allreduce_min_i64 ({sc._peer.transport})" if dp else None),
                 "api": "ShardedCodebook.assign (frame preparation fp32 -> bf16 inside the timed region)"}
    graph = None
    del sc, X
    torch.cuda.empty_cache()
    # ---- VQVAE.quantize with its two 1x1 projections (models/autoencoders.py:142-147), label production
    # (eval, no grad) at the configs[1] quantizer behind 256-channel projections: fused in libpero_b200 against the same
    # module with torch.nn.Conv2d (cuDNN) projections around the VectorQuantizer.  Single GPU only.
    if not dp:
        out["vqvae_quantize"] = vqvae_quantize_leg(dev, flush, steps)
    torch.cuda.empty_cache()
    return out


def roofline_traffic():
    """DRAM bytes per launch of the distance GEMM from the committed ncu capture (profiles/roofline_traffic.json, written by
    tools/summarize_profiles.py together with a hash of the kernel's sources).  A capture taken from other sources is
    not reported: `traffic` is then null instead of a stale number."""
    import hashlib
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu capture committed (profiles/roofline_traffic.json)"
    d = json.load(open(path))
    hsh = hashlib.sha256()
    for f in d.get("sources", []):
        with open(os.path.join(ROOT, f), "rb") as fh:
            hsh.update(fh.read())
    if hsh.hexdigest() != d.get("sources_sha256"):
        return None, f"capture {d.get('capture')} predates the current kernel sources"
    return d["traffic_bytes"], f"{d.get('capture')}: dram read {d['dram_read_bytes'] / 1e6:.2f} MB + write {d['dram_write_bytes'] / 1e6:.2f} MB per launch"


def _log(msg):
    if os.environ.get("PERO_BENCH_VERBOSE"):
        sys.stderr.write(f"[bench rank {os.environ.get('RANK', '0')} t={time.time() % 1000:.2f}] {msg}\n")
        sys.stderr.flush()


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer exists: pinned
    memory is placed on the NUMA node of the allocating thread, and a far node roughly halves the H2D bandwidth the
    e2e leg depends on (and adds latency to every launch).  Returns (all cpus, bound cpus) or (all, None)."""
    all_cpus = sorted(os.sched_getaffinity(0))
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local_rank)
        try:
            bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(all_cpus) // 64) + 1)
        local = [c for c in all_cpus if (int(words[c // 64]) >> (c % 64)) & 1]
        if local and len(local) < len(all_cpus):
            os.sched_setaffinity(0, local)
            return all_cpus, local
    except Exception:
        pass
    return all_cpus, None


def our_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dp = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus, bound_cpus = bind_to_gpu_numa_node(local_rank)
    if dp:
        torch.distributed.init_process_group("nccl", device_id=dev)
    from pero_pretraining_b200 import ops
    ops.require_device()
    c = CFG
    batch = make_batch(rank)
    ds = DeviceStep(batch, dev, dp)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    # ---- warm-up (eager), kernel count, optional CUDA graph
    _log("setup done")
    for _ in range(max(3, args.warmup)):
        ds()
    torch.cuda.synchronize()
    _log("eager warm-up done")
    n_kernels, n_ours = count_kernels(ds, torch.distributed.barrier if dp else None)
    _log(f"kernel count {n_kernels}")
    graph = None
    if not args.no_graph:
        try:
            # The capture stream carries the quantize chain and the distance GEMM: high priority, like the masked-CE
            # stream; only the EMA chain (and the head operand preparation) run at low priority and fill in beside.
            s = torch.cuda.Stream(priority=-1)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                ds()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                ds()
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
        except Exception as e:          # noqa: BLE001 — fall back to eager launches, say so in the config
            graph = None
            sys.stderr.write(f"[bench] CUDA graph capture unavailable ({type(e).__name__}: {e}); eager launches\n")
            torch.cuda.synchronize()
    _log(f"graph {'captured' if graph is not None else 'unavailable'}")
    run = graph.replay if graph is not None else ds
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup + 400):       # ~0.1 s under load before the timed region (a FIXED count: every rank
        run()                                # must issue the same number of collectives); nvidia-smi starts sampling
    torch.cuda.synchronize()

    # ---- timed region: K steps, device time per step, L2 flushed between steps, max over ranks
    torch.cuda.synchronize()
    if dp:
        torch.distributed.barrier()
    t_region0 = time.time()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in ev:
        flush.zero_()
        e0.record()
        run()
        e1.record()
    torch.cuda.synchronize()
    if dp:
        torch.distributed.barrier()
    clocks = sampler.stop(t_region0, time.time())
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    m_total = ds.m_global
    if dp:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = m_total / (ms_per_step * 1e-3)

    if args.timeline and rank == 0:
        dump_timeline(run, args.timeline)
    elif args.timeline:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
    # ---- roofline of the dominant kernel, e2e, CPU baseline
    _log(f"timed region done: {ms_per_step * 1e3:.1f} us/step")
    N = c["lines"] * c["frames"]
    gemm_ms, gemm_min, gemm_sets = gemm_roofline_leg(ds, dev, 20, flush)
    burst, sustained, hbm, src = peaks()
    flops = 2.0 * N * c["K"] * c["D"]
    achieved = flops / (gemm_ms * 1e-3) / 1e12
    traffic, traffic_src = roofline_traffic()
    roofline = {"bound": "tensor", "kernel": "gemm_tn_kernel<2,true,ArgminEpi> (distance GEMM + arg-min)",
                "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": 2.0 * (N * c["D"] + c["K"] * c["D"]) + 4.0 * c["K"] + 8.0 * N,
                "peak_source": f"{src} bf16 burst (kernel timed alone)", "kernel_us": gemm_ms * 1e3, "kernel_us_min": gemm_min * 1e3,
                "timing": (f"{gemm_sets} launches back to back on {gemm_sets} distinct operand sets (> L2 in total) between two CUDA "
                           "events, x20; the launches are programmatic (each waits for its predecessor before its first global "
                           "access), as in a label-production loop"),
                "algorithmic_flops_per_launch": flops,
                "step_tensor_tflops": (flops + 6.0 * ds.M * c["Dh"] * c["V"]) / (ms_per_step * 1e-3) / 1e12,
                "step_frac_of_sustained": (flops + 6.0 * ds.M * c["Dh"] * c["V"]) / (ms_per_step * 1e-3) / 1e12 / sustained}
    _log("roofline leg done")
    e2e = None
    if not args.skip_e2e:
        st_autograd = os.environ.get("PERO_E2E_ST", "1") != "0"
        loss_read = os.environ.get("PERO_E2E_LOSS", "pipelined")
        # Two loops that differ only in how the step's loss reaches the host (both read EVERY step's loss inside the timed
        # region): a blocking `loss.item()` per step, and an asynchronous D2H into pinned memory that the host reads one
        # step later.  Which is faster depends on the host (on a loaded box the blocking loop wins); both medians are
        # reported and the headline is the faster of the two (all values are maxima over the ranks, so every rank picks
        # the same one).
        legs = {}
        for mode in ("sync", "pipelined") if loss_read == "pipelined" else (loss_read,):
            legs[mode] = e2e_leg(batch, dev, dp, args.steps, max(3, args.warmup), loss_read=mode, single_thread_autograd=st_autograd)
        best = min(legs, key=lambda m: legs[m][0])
        s_per_step, s_min, s_all, h2d, d2h, e2e_loss = legs[best]
        describe = {"pipelined": ("every step's loss copied D2H into pinned memory asynchronously and read by the host one step "
                                  "later, the last one before the closing synchronize"),
                    "sync": "blocking loss.item() at the end of every step"}
        e2e = {"value": m_total / s_per_step, "unit": "masked frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": s_per_step * 1e3, "ms_per_step_min": s_min * 1e3, "ms_per_step_all": [t * 1e3 for t in s_all],
               "statistic": f"median of {len(s_all)} timed loops of {args.steps} steps; the faster of the loss read-back modes measured",
               "loss_read": describe[best],
               "loss_read_modes": {m: {"how": describe[m], "ms_per_step": v[0] * 1e3, "ms_per_step_min": v[1] * 1e3,
                                       "ms_per_step_all": [t * 1e3 for t in v[2]], "last_loss": v[5]} for m, v in legs.items()},
               "autograd": "single-threaded (torch.autograd.set_multithreading_enabled(False))" if st_autograd else "default",
               "last_loss": e2e_loss,
               "inputs": "x fp32 [64,256,1,128] + hidden states bf16 [64,128,512] + masked-row list, pinned host memory",
               "api": ("VectorQuantizer.forward (enable_cuda_graph" + (", data parallel" if dp else "") +
                       ") / calculate_loss + LinearHead.masked_loss + backward")}
    _log("e2e leg done")
    # the step's graph, exchange buffers and streams are no longer needed: free them before the large configs
    launch_mode = "cuda_graph" if graph is not None else "eager"
    graph = run = None
    ds_M, ds_info = ds.M, ((ds.peer.transport, ds.peer.n_blocks, len(ds.v_ranges)) if dp else None)
    ds = None
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    gpu_base = None
    if rank == 0 and world == 1 and not args.skip_gpu_baseline:
        gpu_base = gpu_baseline_leg(batch, dev, flush)
        gpu_base["ours_over_fp32"] = value / gpu_base["fp32"]["value"]
        gpu_base["ours_over_bf16_autocast_head"] = value / gpu_base["bf16_autocast_head"]["value"]
    _log("gpu baseline leg done")
    dp_check = dp_equivalence_check(dev, rank, world) if (dp and not args.skip_dp_check) else None
    _log("data-parallel equivalence check done")
    configs = None
    if not args.skip_configs:
        configs = configs_legs(dev, dp, rank, world, flush, max(5, min(args.steps, 20)))
    _log("configs legs done")
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        os.sched_setaffinity(0, all_cpus)              # the CPU arm gets every host core again
        mean_s, min_s, _, _ = run_cpu(make_batch(0), 3, 1)
        cpu = {"value": ds_M / mean_s, "unit": "masked frames/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "3 full-size steps of one 64-line batch after 1 warm-up", "ms_per_step": mean_s * 1e3}

    if rank == 0:
        line = {"metric": "masked_frames_per_sec", "value": value, "unit": "masked frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, **{k: c[k] for k in ("lines", "frames", "K", "D", "Dh", "V", "p")},
                           "masked_frames_per_step": m_total, "frames_per_step": N * world, "parallelism": f"dp{world}" if dp else "single",
                           "exchange": (f"libpero peer all-reduce ({ds_info[0]}, {ds_info[1]} CTAs), head backward in "
                                        f"{ds_info[2]} label ranges" if dp else None),
                           "l2": "flushed (256 MiB write) between timed steps", "launch": launch_mode,
                           "host_affinity": (f"{len(bound_cpus)} CPUs local to the GPU (NVML)" if bound_cpus else "unchanged")},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_baseline": gpu_base, "configs": configs,
                "dp_check": dp_check,
                "gpu_launches": (n_ours if n_ours > 0 else n_kernels) * args.steps, "kernels_per_step": n_kernels, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if dp:
        # A captured graph that contains NCCL collectives must be gone before the communicator is torn down,
        # and a stuck teardown must not keep the job alive after the result line is out.
        torch.cuda.synchronize()
        torch.distributed.barrier()
        sys.stdout.flush()
        threading.Timer(15.0, lambda: os._exit(0)).start()
        try:
            torch.distributed.destroy_process_group()
        finally:
            os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        our_arm(a)
