#!/usr/bin/env python
"""Headline benchmark of the B200-native PP-YOLO hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision f16x2|bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json metric "images/sec at bs=32 608x608", configs[2]): ppyolo_2x (ResNet50-vd + DCNv2 stage 5 +
IoU-aware/CoordConv/SPP head) end-to-end inference -- backbone, head, box decode, Matrix-NMS -- on synthetic N(0,1)
images and seeded random weights, batch 32 per GPU (weak scaling: every rank owns its own 32 images end to end, no
collective on the data path).

The headline runs the engine's DEFAULT precision 'f16x2': the fp32-grade tensor-core mode (fp16 hi/lo pair operands,
three tcgen05 MMAs per K block, K-chunked fp32 accumulation) whose detections are parity-tested against the fp32 reference
at detection level (tests/test_gpu_engine.py).  The bf16 mode is reported beside it as `precision_modes.bf16`, with its
measured drift from the headline mode's detections on the same batch -- it is NOT the headline.

One step = one forward of one batch.  `value` is device-timed (CUDA events, max over ranks) with the input batch already
in HBM; `e2e` is the same metric through the public PPYOLO/engine API with the inputs in pinned HOST memory (upload of batch
i+1 overlaps the compute of batch i; detections + counts come back every step).

Sub-records of the same JSON line:
  roofline            conv kernel family of one step, timed live as its own CUDA graph; peak = measured dense bf16/fp16
                      tensor peak (burst or sustained by the clocks seen while timing) / 3 MMAs per algorithmic MAC
  cpu_baseline        oracle port of the reference forward on the host cores, bounded sample (bs 4)
  torch_gpu_baseline  the same restatement in stock PyTorch (cuDNN/cuBLAS) on THIS GPU: fp32 (TF32 off), TF32, bf16
                      channels_last -- the reference's real GPU path is stock PyTorch; backbone + head + decode (no NMS loop)
  train               BASELINE configs[3] (ppyolo_2x 608^2, bs 8/GPU train step with NCCL gradient all-reduce), at every N
  matrix_nms          BASELINE configs[4] (10k boxes x 80 classes) in isolation, us per image

`--impl reference` times the reference's CPU path (oracle port of model/ppyolo.py:19-22 on the host cores; the Python
reference itself cannot travel to the GPU box) on a bounded sample of the same workload; it maps no native code of this repo.
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200'))
sys.path.insert(0, REPO)

import torch  # noqa: E402

ARCH, SIZE, BATCH = 'r50vd', 608, 32
WORKLOAD = 'ppyolo_2x ResNet50-vd+DCNv2 608x608 bs=32/GPU inference (backbone+head+decode+MatrixNMS)'
METRIC, UNIT = 'images_per_sec', 'images/s'
DTYPE_LABEL = {'f16x2': 'fp32-grade (f16x2: fp16 hi/lo pair operands, 3 tcgen05 MMAs per K block, fp32 accumulate)',
               'bf16': 'bf16 (tcgen05, fp32 accumulate)', 'fp32': 'f32 (SIMT)'}


def build_model(arch, train=False):
    """Parameter containers + seeded weights.  Loads no native code (the kernel library is mapped on first use)."""
    import config as cfgs
    from model.ppyolo import PPYOLO
    from ppyolo_b200 import synth
    cfg = {'r50vd': cfgs.PPYOLO_2x_Config, 'r18vd': cfgs.PPYOLO_r18vd_Config}[arch]()
    backbone = cfgs.select_backbone(cfg.backbone_type)(**cfg.backbone)
    yolo = None
    if train:
        iou_loss = cfgs.select_loss(cfg.iou_loss_type)(**cfg.iou_loss)
        iou_aware = cfgs.select_loss(cfg.iou_aware_loss_type)(**cfg.iou_aware_loss) if cfg.head['iou_aware'] else None
        yolo = cfgs.select_loss(cfg.yolo_loss_type)(iou_loss=iou_loss, iou_aware_loss=iou_aware, **cfg.yolo_loss)
    head = cfgs.select_head(cfg.head_type)(yolo_loss=yolo, is_train=train, nms_cfg=cfg.nms_cfg, **cfg.head)
    model = PPYOLO(backbone, head)
    synth.randomize_(model, seed=0)
    if train:
        model.train()
        backbone.freeze()
    else:
        model.eval()
        head.set_dropblock(is_test=True)
    return model, cfg


def measured_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs (NVML is initialised in
    the constructor so the first sample falls inside the region; 5 ms period)."""
    NAMES = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40,
             'hw_power_brake': 0x80}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self._nvml, self._handle = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # NVML missing: report that, do not fail the bench
            self.reasons.add('nvml_unavailable:%s' % type(exc).__name__)

    def _sample(self):
        nv, h = self._nvml, self._handle
        self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for k, bit in self.NAMES.items():
            if r & bit:
                self.reasons.add(k)

    def run(self):
        if self._handle is None:
            return
        try:
            while not self._halt.is_set():
                self._sample()
                time.sleep(0.005)
        except Exception as exc:
            self.reasons.add('nvml_error:%s' % type(exc).__name__)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


# ------------------------------------------------------------------------------------------------
# reference arms: the oracle port on the host cores / in stock PyTorch on the GPU
# ------------------------------------------------------------------------------------------------
CPU_SAMPLE_BATCH = 4


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


def cpu_reference_rate(steps, warmup, batch=CPU_SAMPLE_BATCH, arch=ARCH, size=SIZE):
    """Oracle port of the reference forward on the host cores: images/s on a bounded sample (``batch`` images per step,
    the reference's own eval_batch_size is 4: config/ppyolo_2x.py:77)."""
    from oracle import ppyolo_ref as ref
    from ppyolo_b200 import synth
    model, cfg = build_model(arch)
    net = ref.Net(model.state_dict(), cfg)
    cores = torch.get_num_threads()
    x = synth.images(batch, size, seed=1)
    im = synth.im_sizes(batch)
    for _ in range(warmup):
        net.forward(x, im)
    t0 = time.perf_counter()
    for _ in range(steps):
        net.forward(x, im)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt * 1e3, cores


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 50)), max(1, min(args.warmup, 3))      # ~0.8 s per 4-image step: K = 20, W = 5 is ~20 s
    torch.set_num_threads(host_threads())          # torchrun pins OMP_NUM_THREADS=1; the reference arm may use every host core
    rate, ms, cores = cpu_reference_rate(steps, warmup)
    sample = ('%d forward(s) of %d images 608x608 (a slice of the 32-image batch), backbone+head+decode+MatrixNMS, torch CPU '
              'fp32, %d threads' % (steps, CPU_SAMPLE_BATCH, cores))
    from ppyolo_b200 import _lib
    line = {'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'sample': sample},
            'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'native_library_loaded': _lib.is_loaded()}
    print(json.dumps(line), flush=True)


def torch_gpu_baseline(dev, steps=5, warmup=3):
    """The reference's REAL GPU path is stock PyTorch ("the honest competitor"): the oracle restatement of
    model/ppyolo.py:19-22 run on this GPU with cuDNN/cuBLAS kernels only -- backbone + head + box decode at bs 32 x 608^2 (the
    per-image Matrix-NMS loop left out, in the baseline's favour), CUDA events, three arithmetic variants."""
    from oracle import ppyolo_ref as ref
    from ppyolo_b200 import synth
    model, cfg = build_model(ARCH)
    sd = model.state_dict()
    x = synth.images(BATCH, SIZE, seed=10).to(dev)
    im = synth.im_sizes(BATCH).to(dev)
    out = {'sample': 'bs %d x %d^2, backbone+head+decode (no NMS), %d timed steps after %d warm-ups, CUDA events' % (BATCH, SIZE, steps, warmup),
           'torch': torch.__version__, 'cudnn': torch.backends.cudnn.version()}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    variants = (('fp32_tf32_off', False, torch.float32, False), ('tf32', True, torch.float32, False),
                ('bf16_channels_last', True, torch.bfloat16, True))
    try:
        torch.backends.cudnn.benchmark = True
        for name, tf32, dtype, cl in variants:
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            try:
                net = ref.Net(sd, cfg, device=dev, dtype=dtype, channels_last=cl)
                for _ in range(warmup):
                    net.forward_dense(x, im)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    net.forward_dense(x, im)
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1) / steps
                out[name] = {'value': BATCH / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms}
                del net
            except Exception as exc:      # a variant cuDNN refuses must not sink the bench line
                out[name] = {'error': '%s: %s' % (type(exc).__name__, str(exc)[:200])}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return out


def matrix_nms_isolation(dev, with_cpu):
    """BASELINE configs[4] / metric "MatrixNMS us/img": 10k synthetic boxes x 80 classes per image through the C-ABI batched
    Matrix-NMS (device-resident inputs, CUDA events), at bs 32 and bs 1; next to it the oracle port of model/matrix_nms.py on
    one image on the host."""
    from ppyolo_b200 import ops, synth
    b, s = synth.nms_inputs(10000, 80, seed=0)
    res = {'config': '10k boxes x 80 classes per image, score_thr=post_thr=0.01, top_k=500, keep=100 (configs[4])'}
    for bs in (32, 1):
        boxes = b[None].repeat(bs, 1, 1).to(dev).contiguous()
        scores = s[None].repeat(bs, 1, 1).to(dev).contiguous()
        out = torch.empty((bs, 100, 6), dtype=torch.float32, device=dev)
        counts = torch.empty((bs,), dtype=torch.int32, device=dev)
        ws = ops.nms_workspace(bs, 10000, 80, dev)
        run = lambda: ops.matrix_nms_launch(boxes, scores, out, counts, ws, 0.01, 0.01, 500, 100, False, 2.0)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        res['us_per_img_bs%d' % bs] = e0.elapsed_time(e1) / 20 / bs * 1e3
    if with_cpu:
        from oracle import ppyolo_ref as ref
        t0 = time.perf_counter()
        ref.matrix_nms(b.numpy(), s.numpy(), 0.01, 0.01, 500, 100)
        res['cpu_port_us_per_img'] = (time.perf_counter() - t0) * 1e6
    return res


def is_glue(name):
    """Plan steps that are not the conv kernel family (the DCN sampling stage `.gather` belongs to the family: its GEMM's
    FLOPs are counted, so its time is too)."""
    return name.startswith(('nchw', 'maxpool', 'avgpool', 'spp', 'decode', 'matrix_nms', 'reset_flags')) or name == 'stem.conv1_1'


def per_op_times(eng, iters=3):
    """CUDA-event duration of every plan step (eager, on the launching stream), averaged over `iters`."""
    stream = torch.cuda.current_stream()
    acc = [0.0] * len(eng.steps)
    for _ in range(iters):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(eng.steps) + 1)]
        evs[0].record(stream)
        for i, (_, fn) in enumerate(eng.steps):
            fn()
            evs[i + 1].record(stream)
        torch.cuda.synchronize()
        for i in range(len(eng.steps)):
            acc[i] += evs[i].elapsed_time(evs[i + 1]) / iters
    return [(eng.steps[i][0], acc[i]) for i in range(len(eng.steps))]


def detection_drift(got, want, iou_thr=0.5):
    """Fraction of `want` rows matched by a row of `got` with the same label and IoU >= iou_thr, label agreement at equal
    rank, and the largest box / score difference over the matched pairs (numpy [M,6] rows per image)."""
    import numpy as np
    total = matched = same_rank = 0
    max_box = max_score = 0.0
    for g, w in zip(got, want):
        g, w = np.asarray(g, dtype=np.float64), np.asarray(w, dtype=np.float64)
        g = g[:0] if (g.shape[0] == 1 and g[0, 0] < 0) else g
        w = w[:0] if (w.shape[0] == 1 and w[0, 0] < 0) else w
        total += len(w)
        k = min(len(g), len(w))
        same_rank += int((g[:k, 0] == w[:k, 0]).sum())
        used = np.zeros(len(g), dtype=bool)
        for row in w:
            cand = np.where((g[:, 0] == row[0]) & ~used)[0]
            if not len(cand):
                continue
            b = g[cand, 2:]
            ix = np.clip(np.minimum(b[:, 2], row[4]) - np.maximum(b[:, 0], row[2]), 0, None)
            iy = np.clip(np.minimum(b[:, 3], row[5]) - np.maximum(b[:, 1], row[3]), 0, None)
            inter = ix * iy
            union = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) + (row[4] - row[2]) * (row[5] - row[3]) - inter
            iou = inter / np.maximum(union, 1e-12)
            j = int(np.argmax(iou))
            if iou[j] >= iou_thr:
                used[cand[j]] = True
                matched += 1
                max_box = max(max_box, float(np.abs(b[j] - row[2:]).max()))
                max_score = max(max_score, abs(float(g[cand[j], 1] - row[1])))
    return {'reference_rows': total, 'matched_frac_iou%.2f' % iou_thr: matched / max(total, 1),
            'same_label_at_rank_frac': same_rank / max(total, 1), 'max_box_err_px_matched': max_box, 'max_score_err_matched': max_score}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Dist(object):
    def __init__(self, world, dev):
        self.world, self.dev = world, dev

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, ms):
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms


def time_engine(eng, steps, warmup, dd, sample_clocks=None):
    main = torch.cuda.current_stream()
    for _ in range(warmup):
        eng.launch()
    dd.barrier()
    sampler = ClockSampler(sample_clocks) if sample_clocks is not None else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(steps):
        eng.launch()
    e1.record(main)
    dd.barrier()
    clocks = sampler.stop() if sampler else None
    return dd.max(e0.elapsed_time(e1)) / steps, clocks


def time_e2e(eng, x_host, im_host, steps, warmup, dd, dev):
    """Public-API path with HOST buffers: two input slots (one captured graph each); batch i+1 is uploaded straight into the
    engine on a copy stream while batch i computes; detections and counts are downloaded every step."""
    main = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream(device=dev)
    while len(eng.input_slots) < 2:
        eng.add_input_slot()
    out_host = [torch.empty((eng.n, eng.keep_top_k, 6), dtype=torch.float32).pin_memory() for _ in range(2)]
    cnt_host = [torch.empty((eng.n + 1,), dtype=torch.int32).pin_memory() for _ in range(2)]
    stage_im = [torch.empty_like(eng.im_size) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def loop(n_steps):
        for b in range(2):
            consumed[b].record(main)
        for step in range(n_steps):
            b = step % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])
                eng.input_slots[b].copy_(x_host[b], non_blocking=True)
                stage_im[b].copy_(im_host[b], non_blocking=True)
                ready[b].record(copy_stream)
            main.wait_event(ready[b])
            eng.im_size.copy_(stage_im[b], non_blocking=True)
            eng.launch(slot=b)
            consumed[b].record(main)
            out_host[b].copy_(eng.nms_out, non_blocking=True)
            cnt_host[b].copy_(eng._flags, non_blocking=True)
            done[b].record(main)
            if step >= 1:
                done[1 - b].synchronize()        # the caller consumes the previous step's detections
                _ = int(cnt_host[1 - b][0])
        done[(n_steps - 1) % 2].synchronize()

    loop(max(2, warmup))
    dd.barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(main)
    loop(steps)
    s1.record(main)
    dd.barrier()
    ms = dd.max(s0.elapsed_time(s1)) / steps
    h2d = x_host[0].numel() * x_host[0].element_size() + im_host[0].numel() * 4
    d2h = out_host[0].numel() * 4 + cnt_host[0].numel() * 4
    return ms, h2d, d2h


def time_e2e_raw(eng, blobs, metas, im_host, steps, warmup, dd, dev):
    """End to end from the ORIGINAL images: packed uint8 BGR images of the camera / file size in pinned host
    memory -> one upload -> batched bicubic resize + BGR->RGB kernel straight into the engine's uint8 input slot -> forward ->
    detections back.  The host does no image processing at all."""
    from ppyolo_b200 import ops
    main = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream(device=dev)
    while len(eng.input_slots) < 2:
        eng.add_input_slot()
    blob_dev = [torch.empty_like(b, device=dev) for b in blobs]
    meta_dev = [m.to(dev) for m in metas]
    out_host = [torch.empty((eng.n, eng.keep_top_k, 6), dtype=torch.float32).pin_memory() for _ in range(2)]
    cnt_host = [torch.empty((eng.n + 1,), dtype=torch.int32).pin_memory() for _ in range(2)]
    stage_im = [torch.empty_like(eng.im_size) for _ in range(2)]
    ready, consumed, done = ([torch.cuda.Event() for _ in range(2)] for _ in range(3))

    def loop(n_steps):
        for b in range(2):
            consumed[b].record(main)
        for step in range(n_steps):
            b = step % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])
                blob_dev[b].copy_(blobs[b], non_blocking=True)
                stage_im[b].copy_(im_host[b], non_blocking=True)
                ready[b].record(copy_stream)
            main.wait_event(ready[b])
            ops.resize_cubic_u8(blob_dev[b], meta_dev[b], eng.h, swap_rb=True, out=eng.input_slots[b])
            eng.im_size.copy_(stage_im[b], non_blocking=True)
            eng.launch(slot=b)
            consumed[b].record(main)
            out_host[b].copy_(eng.nms_out, non_blocking=True)
            cnt_host[b].copy_(eng._flags, non_blocking=True)
            done[b].record(main)
            if step >= 1:
                done[1 - b].synchronize()
                _ = int(cnt_host[1 - b][0])
        done[(n_steps - 1) % 2].synchronize()

    loop(max(2, warmup))
    dd.barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(main)
    loop(steps)
    s1.record(main)
    dd.barrier()
    ms = dd.max(s0.elapsed_time(s1)) / steps
    return ms, blobs[0].numel() + metas[0].numel() * 8 + im_host[0].numel() * 4


def conv_family_roofline(eng, precision, steps, ms_step, local_rank):
    """Roofline of the dominant kernel family (tcgen05 implicit-GEMM conv, all launches of one step), rank 0."""
    peaks, peak_src = measured_peaks()
    main = torch.cuda.current_stream()
    ops_t = per_op_times(eng, iters=3)
    conv_ms_eager = sum(t for name, t in ops_t if not is_glue(name))
    conv_launches = sum(1 for name, _ in ops_t if not is_glue(name))
    # the conv launches of one step replayed back to back as their own CUDA graph (same parameters and buffers, glue kernels
    # left out): the kernel family under the conditions of the timed region -- graph replay with programmatic dependent launch
    conv_graph = eng.capture_steps(lambda name: not is_glue(name))
    for _ in range(3):
        conv_graph.replay()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(main)
    for _ in range(steps):
        conv_graph.replay()
    c1.record(main)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    conv_ms = c0.elapsed_time(c1) / steps
    achieved = eng.conv_flops / (conv_ms * 1e-3) / 1e12
    # which measured peak: the burst figure when the clocks stayed near the maximum while timing, the sustained one otherwise
    burst = bool(clocks['sm_mhz'] and clocks['sm_max_mhz'] and clocks['sm_mhz'] >= 0.93 * clocks['sm_max_mhz'])
    dense = peaks['bf16_tflops'] if burst else peaks['bf16_tflops_sustained']
    mma_per_mac = 3 if precision == 'f16x2' else 1
    peak = dense / mma_per_mac if precision in ('bf16', 'f16x2') else 75.0
    roofline = {'bound': 'tensor', 'kernel': 'conv_umma_kernel (all %d conv launches of one step)' % conv_launches,
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None,
                'peak_source': '%s %s (SM clock median %s of %s MHz while timing) / %d MMAs per algorithmic MAC' % (
                    peak_src, 'bf16_tflops [burst]' if burst else 'bf16_tflops_sustained', clocks['sm_mhz'], clocks['sm_max_mhz'], mma_per_mac),
                'frac_of_burst_peak': achieved * mma_per_mac / peaks['bf16_tflops'],
                'frac_of_sustained_peak': achieved * mma_per_mac / peaks['bf16_tflops_sustained'],
                'achieved_mma_tflops': achieved * mma_per_mac,
                'min_bytes_per_step': sum(v['bytes'] for k, v in eng.step_info.items() if not is_glue(k)),
                'flops_per_step': eng.conv_flops, 'kernel_ms_per_step': conv_ms, 'share_of_step': conv_ms / ms_step,
                'timing': 'CUDA events around %d replays of a CUDA graph holding the %d conv launches of one step' % (steps, conv_launches),
                'kernel_ms_per_step_eager_events': conv_ms_eager, 'achieved_eager_events': eng.conv_flops / (conv_ms_eager * 1e-3) / 1e12,
                'traffic_note': 'see profiles/ (ncu launch list of this round) for DRAM bytes per launch'}
    # DRAM bytes of the same launches from the committed ncu launch list of this round (profiles/, tools/summarize_launches.py)
    tpath = os.path.join(REPO, 'profiles', 'r02_conv_family_traffic_%s.json' % precision)
    if os.path.exists(tpath):
        with open(tpath) as f:
            t = json.load(f)
        roofline['traffic'] = t['dram_bytes']
        roofline['traffic_note'] = ('dram__bytes_read.sum + dram__bytes_write.sum over the %d conv launches of one step, ncu launch list '
                                    'profiles/r02_ncu_launches_%s.csv; algorithmic minimum = min_bytes_per_step' % (t['launches'], precision))
    os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(REPO, 'gpurun_out', 'per_op_ms_%s.json' % precision), 'w') as f:
        json.dump({'ops': ops_t, 'total_ms': sum(t for _, t in ops_t), 'graph_ms_per_step': ms_step, 'info': eng.step_info}, f, indent=1)
    return roofline


def train_record(args, rank, world, local_rank, dd, dev):
    """BASELINE configs[3]: ppyolo_2x 608x608, bs 8 per GPU, train.py step -- frozen-backbone forward with batch-stat BN, head
    forward + backward, six losses, NCCL all-reduce of the 92.6 MB gradient bucket, fused SGD -- on synthetic data."""
    from ppyolo_b200 import synth, targets as tg
    from ppyolo_b200.trainer import Trainer
    bs = 8
    model, cfg = build_model(ARCH, train=True)
    model = model.to(dev)
    model.train_precision = 'bf16'
    trainer = Trainer(model, cfg, graph=True)
    x = synth.images(bs, SIZE, seed=20 + rank).to(dev)
    gb, gc, gs = tg.synthetic_ground_truth(bs, seed=30 + rank)
    targets = [torch.from_numpy(t).to(dev) for t in tg.gt2yolo_target(gb, gc, gs, h=SIZE, w=SIZE, **cfg.gt2YoloTarget)]
    gb, gc, gs = (torch.from_numpy(v).to(dev) for v in (gb, gc, gs))
    steps, warmup = max(5, min(args.steps, 20)), 3
    for _ in range(warmup):
        losses = trainer.step(x, gb, gc, gs, targets)
    dd.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        losses = trainer.step(x, gb, gc, gs, targets)
    e1.record()
    dd.barrier()
    ms = dd.max(e0.elapsed_time(e1)) / steps
    rec = {'workload': 'ppyolo_2x 608x608 bs=8/GPU train step (freeze_at=5: frozen backbone fwd with batch-stat BN, head fwd+bwd, 6 losses, '
                       'gradient exchange + SGD + EMA)', 'value': world * bs / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps,
           'warmup': warmup, 'n_gpus': world, 'scaling': 'weak', 'backbone_precision': 'bf16',
           'allreduce_bytes': int(trainer.bucket.flat.numel() * 4), 'trainable_params': int(sum(p.numel() for p in trainer.params)),
           'losses': {k: float(v) for k, v in losses.items()}}
    if hasattr(trainer, 'timing_summary'):
        ts = trainer.timing_summary()
        if world > 1:
            # the exchange step rendezvouses the ranks, so its event time on a rank includes waiting for the slowest one: report the
            # fastest rank's view (the work itself) next to this rank's
            import torch.distributed as dist
            for key in ('allreduce_ms', 'exchange_optimizer_ms'):
                if ts.get(key) is not None:
                    t = torch.tensor([ts[key]], dtype=torch.float32, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MIN)
                    ts[key + '_rank0_incl_wait'] = ts[key]
                    ts[key] = float(t.item())
        rec.update(ts)
    del trainer, model
    torch.cuda.empty_cache()
    return rec


def run_ours(args, rank, world, local_rank):
    from ppyolo_b200 import synth, _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dd = Dist(world, dev)
    model, cfg = build_model(ARCH)
    model = model.to(dev)
    model.precision = args.precision
    eng = model.engine(BATCH, SIZE, SIZE)
    x_host = [synth.images(BATCH, SIZE, seed=10 + rank * 2 + i).pin_memory() for i in range(2)]
    im_host = [synth.im_sizes(BATCH).pin_memory() for _ in range(2)]
    eng.x_in.copy_(x_host[0])
    eng.im_size.copy_(im_host[0])

    # ---- device-resident throughput, then end to end through the public API with host buffers
    ms_step, clocks = time_engine(eng, args.steps, args.warmup, dd, sample_clocks=local_rank)
    value = world * BATCH / (ms_step * 1e-3)
    flags = eng._flags.cpu()
    assert int(flags[:BATCH].min()) >= 0, 'matrix_nms overflow flag'
    assert int(flags[BATCH]) == 0, 'f16x2 activation overflow flag'
    cand_mean = float(eng.candidate_counts().float().mean()) if eng.scores is None else None
    head_preds = [p.cpu().numpy() for p in model(x_host[0].to(dev), im_host[0].to(dev))]
    # e2e: the public fast path -- resized uint8 RGB batches in pinned host memory (Decode.process_image_u8), NormalizeImage +
    # Permute inside the stem kernel: 35.5 MB per step instead of the 141.9 MB of the float CHW tensor (also timed, for reference)
    u8_host = [synth.images_u8(BATCH, SIZE, seed=10 + rank * 2 + i).pin_memory() for i in range(2)]
    eng_u8 = model.engine(BATCH, SIZE, SIZE, input_u8=True)
    e2e_ms, h2d, d2h = time_e2e(eng_u8, u8_host, im_host, args.steps, args.warmup, dd, dev)
    e2e_f32_ms, h2d_f32, _ = time_e2e(eng, x_host, im_host, args.steps, args.warmup, dd, dev)
    # the same from the ORIGINAL images: 480x640 BGR uint8 frames packed in pinned memory, resized on the GPU (csrc/preprocess.cu)
    from ppyolo_b200 import ops
    import numpy as np
    raw_rng = np.random.RandomState(100 + rank)
    packed = [ops.pack_images([raw_rng.randint(0, 256, (480, 640, 3)).astype(np.uint8) for _ in range(BATCH)]) for _ in range(2)]
    e2e_raw_ms, h2d_raw = time_e2e_raw(eng_u8, [p[0] for p in packed], [p[1] for p in packed], im_host, args.steps, args.warmup, dd, dev)
    del eng_u8, packed

    # ---- the bf16 mode beside it (labelled, with its drift from the headline mode's detections on the same batch)
    modes = {}
    if args.precision == 'f16x2' and not args.quick:
        model.precision = 'bf16'
        eng_b = model.engine(BATCH, SIZE, SIZE)
        eng_b.x_in.copy_(x_host[0])
        eng_b.im_size.copy_(im_host[0])
        ms_b, _ = time_engine(eng_b, args.steps, args.warmup, dd)
        e2e_b, _, _ = time_e2e(eng_b, x_host, im_host, args.steps, args.warmup, dd, dev)
        preds_b = [p.cpu().numpy() for p in model(x_host[0].to(dev), im_host[0].to(dev))]
        modes['bf16'] = {'dtype': DTYPE_LABEL['bf16'], 'value': world * BATCH / (ms_b * 1e-3), 'unit': UNIT, 'ms_per_step': ms_b,
                         'e2e_value': world * BATCH / (e2e_b * 1e-3), 'dcn_impl': eng_b.dcn_impl,
                         'drift_vs_headline_detections': detection_drift(preds_b, head_preds, 0.5),
                         'note': 'throughput mode; NOT parity-grade (8-bit operands through ~55 layers of a random-weight net)'}
        if rank == 0:
            modes['bf16']['roofline'] = conv_family_roofline(eng_b, 'bf16', args.steps, ms_b, local_rank)
        del eng_b
        model.precision = args.precision

    # ---- the north_star's second resolution: the same net and batch at 320x320 (device-resident, CUDA-graph replay)
    shape320 = None
    if not args.quick:
        eng_3 = model.engine(BATCH, 320, 320)
        eng_3.x_in.copy_(synth.images(BATCH, 320, seed=40 + rank))
        eng_3.im_size.copy_(im_host[0])
        ms_3, _ = time_engine(eng_3, args.steps, args.warmup, dd)
        shape320 = {'workload': 'ppyolo_2x 320x320 bs=%d/GPU inference' % BATCH, 'value': world * BATCH / (ms_3 * 1e-3), 'unit': UNIT,
                    'ms_per_step': ms_3, 'conv_gflop_per_image': eng_3.conv_flops / BATCH / 1e9}
        del eng_3

    # ---- strong scaling: the SAME 32 images split over the ranks, 32 / N per GPU -- the harder number (small M:
    # tile quantisation); N > 1 only, N = 1 is the headline itself
    strong = None
    if world > 1 and BATCH % world == 0 and not args.quick:
        bs = BATCH // world
        eng_s = model.engine(bs, SIZE, SIZE)
        eng_s.x_in.copy_(x_host[0][:bs])
        eng_s.im_size.copy_(im_host[0][:bs])
        ms_s, _ = time_engine(eng_s, args.steps, args.warmup, dd)
        strong = {'scaling': 'strong', 'global_batch': BATCH, 'batch_per_gpu': bs, 'value': BATCH / (ms_s * 1e-3), 'unit': UNIT,
                  'ms_per_step': ms_s, 'note': 'device-resident inputs, CUDA-graph replay, max over ranks'}
        del eng_s

    train = None
    if not args.no_train and not args.quick:
        try:
            train = train_record(args, rank, world, local_rank, dd, dev)
        except Exception as exc:      # the training sub-record must not sink the headline line
            train = {'error': '%s: %s' % (type(exc).__name__, str(exc)[:300])}

    if rank != 0:
        return
    roofline = conv_family_roofline(eng, args.precision, args.steps, ms_step, local_rank)
    cpu, tgb = None, None
    if world == 1 and not args.no_cpu_baseline and not args.quick:
        torch.set_num_threads(host_threads())
        rate, ms, cores = cpu_reference_rate(2, 1)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': '2 forwards of %d images 608x608 (a slice of the 32-image batch), torch CPU fp32, %d threads' % (CPU_SAMPLE_BATCH, cores)}
        tgb = torch_gpu_baseline(dev)

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': DTYPE_LABEL[args.precision], 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': world * BATCH, 'precision': args.precision,
                       'l2_policy': 'inputs_exceed_l2 (141.9 MB batch, multi-GB activations per step vs 126 MB L2)',
                       'cuda_graph': True, 'postprocess': eng.postprocess_impl, 'dcn_impl': eng.dcn_impl,
                       'nms_candidates_per_image': cand_mean, 'detections_per_image': float(flags[:BATCH].float().mean()),
                       'sharding': 'batch-sharded replicas, no collective'},
            'clocks': clocks, 'gpu_launches': eng.launches_per_run * args.steps * world,
            'e2e': {'value': world * BATCH / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms, 'h2d_bytes_per_step': world * h2d,
                    'd2h_bytes_per_step': world * d2h, 'input': 'resized uint8 HWC batch (normalise + permute fused into the stem kernel)',
                    'float_chw_upload': {'value': world * BATCH / (e2e_f32_ms * 1e-3), 'ms_per_step': e2e_f32_ms, 'h2d_bytes_per_step': world * h2d_f32},
                    'from_original_images': {'value': world * BATCH / (e2e_raw_ms * 1e-3), 'ms_per_step': e2e_raw_ms, 'h2d_bytes_per_step': world * h2d_raw,
                                             'input': '480x640 BGR uint8 frames, bicubic resize + BGR->RGB on the GPU (no host image processing)'}},
            'roofline': roofline, 'cpu_baseline': cpu, 'torch_gpu_baseline': tgb, 'precision_modes': modes, 'shape_320': shape320, 'strong_scaling': strong, 'train': train,
            'matrix_nms': matrix_nms_isolation(dev, world == 1 and not args.no_cpu_baseline and not args.quick),
            'loaded_library': _lib.LIB_PATH}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='f16x2', choices=['f16x2', 'bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--quick', action='store_true', help='headline + roofline only (development runs)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
