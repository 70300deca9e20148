#!/usr/bin/env python
"""Headline benchmark of the B200-native PP-YOLO hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json metric "images/sec at bs=32 608x608", configs[2]): ppyolo_2x (ResNet50-vd + DCNv2
stage 5 + IoU-aware/CoordConv/SPP head) end-to-end inference -- backbone, head, box decode, Matrix-NMS --
on synthetic N(0,1) images and seeded random weights, batch 32 per GPU (weak scaling: every rank owns its own
32 images end to end, no collective on the data path).

One step = one forward of one batch.  `value` is device-timed (CUDA events, max over ranks) with the input
batch already in HBM; `e2e` is the same metric through the public PPYOLO/engine API with the inputs in pinned
HOST memory: every step uploads its 141.9 MB fp32 batch straight into one of the engine's two input slots (copy
stream; the upload of batch i+1 overlaps the compute of batch i) and downloads the [32,100,6] detections + counts.

`roofline`: the conv kernel family (84 launches of one step) timed live as its own CUDA graph; `traffic` = DRAM bytes of
those launches from the committed ncu launch list.  `matrix_nms`: BASELINE configs[4] (10k boxes x 80 classes) in
isolation, us per image.  `cpu_baseline`: the oracle port of the reference forward on the host cores (bounded sample).

`--impl reference` times the reference's CPU path (oracle port of model/ppyolo.py:19-22 on the host cores;
the Python reference itself cannot travel to the GPU box) on a bounded sample of the same workload.
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


def build_model(arch):
    import config as cfgs
    from model.ppyolo import PPYOLO
    from ppyolo_b200 import synth
    cfg = {'r50vd': cfgs.PPYOLO_2x_Config, 'r18vd': cfgs.PPYOLO_r18vd_Config}[arch]()
    backbone = cfgs.select_backbone(cfg.backbone_type)(**cfg.backbone)
    head = cfgs.select_head(cfg.head_type)(yolo_loss=None, nms_cfg=cfg.nms_cfg, **cfg.head)
    model = PPYOLO(backbone, head)
    synth.randomize_(model, seed=0)
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


def cpu_reference_rate(steps, warmup, arch=ARCH, size=SIZE):
    """Oracle port of the reference forward on the host cores: images/s on a bounded sample (1 image/step)."""
    from oracle import ppyolo_ref as ref
    from ppyolo_b200 import synth
    model, cfg = build_model(arch)
    net = ref.Net(model.state_dict(), cfg)
    cores = torch.get_num_threads()
    x = synth.images(1, size, seed=1)
    im = synth.im_sizes(1)
    for _ in range(warmup):
        net.forward(x, im)
    t0 = time.perf_counter()
    for _ in range(steps):
        net.forward(x, im)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return 1.0 / dt, dt * 1e3, cores


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    try:      # torchrun pins OMP_NUM_THREADS=1; the reference arm may use every host core it is allowed
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    rate, ms, cores = cpu_reference_rate(steps, warmup)
    sample = '%d forward(s) of 1 image 608x608 (of the 32-image batch), torch CPU fp32, %d threads' % (steps, cores)
    line = {'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'sample': sample},
            'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def ncu_conv_traffic():
    """DRAM bytes (read + write) of all conv_umma launches of one step, from the committed ncu launch list of the same
    workload (profiles/r01_ncu_launches.csv, `tools/gpu_profile.sh`); None when the file is missing."""
    import csv
    path = os.path.join(REPO, 'profiles', 'r01_ncu_launches.csv')
    if not os.path.exists(path):
        return None
    total = 0.0
    with open(path) as f:
        rows = list(csv.reader(f))
    try:
        hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
        ix = {h: i for i, h in enumerate(rows[hi])}
        for r in rows[hi + 1:]:
            if len(r) == len(rows[hi]) and 'conv_umma_kernel' in r[ix['Kernel Name']] and \
                    r[ix['Metric Name']] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                total += float(r[ix['Metric Value']].replace(',', ''))
    except (IndexError, KeyError, ValueError):
        return None
    return total or None


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
    return (name.startswith(('nchw', 'maxpool', 'avgpool', 'spp', 'decode', 'matrix_nms')) or name.endswith('.gather')
            or name == 'stem.conv1_1')


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


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from ppyolo_b200 import synth, _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    model, cfg = build_model(ARCH)
    model = model.to(dev)
    model.precision = args.precision
    eng = model.engine(BATCH, SIZE, SIZE)
    x_host = [synth.images(BATCH, SIZE, seed=10 + rank * 2 + i).pin_memory() for i in range(2)]
    im_host = [synth.im_sizes(BATCH).pin_memory() for _ in range(2)]
    out_host = [torch.empty((BATCH, eng.keep_top_k, 6), dtype=torch.float32).pin_memory() for _ in range(2)]
    cnt_host = [torch.empty((BATCH,), dtype=torch.int32).pin_memory() for _ in range(2)]
    eng.x_in.copy_(x_host[0])
    eng.im_size.copy_(im_host[0])
    main = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident throughput -----------------------------------------------------------
    for _ in range(args.warmup):
        eng.launch()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(args.steps):
        eng.launch()
    e1.record(main)
    barrier()
    clocks = sampler.stop()
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * BATCH / (ms_step * 1e-3)
    counts = eng.nms_counts.cpu()
    assert int(counts.min()) >= 0, 'matrix_nms overflow flag'
    cand_mean = float(eng.candidate_counts().float().mean()) if eng.scores is None else None

    # ---- end to end through the public API with host buffers ----------------------------------
    # two input slots (one captured graph each): batch i+1 is uploaded straight into the engine while batch i computes
    copy_stream = torch.cuda.Stream(device=dev)
    if len(eng.input_slots) < 2:
        eng.add_input_slot()
    stage_im = [torch.empty_like(eng.im_size) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(steps):
        for b in range(2):
            consumed[b].record(main)
        for step in range(steps):
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
            cnt_host[b].copy_(eng.nms_counts, non_blocking=True)
            done[b].record(main)
            if step >= 1:
                done[1 - b].synchronize()        # the caller consumes the previous step's detections
                _ = int(cnt_host[1 - b][0])
        done[(steps - 1) % 2].synchronize()

    e2e_loop(max(2, args.warmup))
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(main)
    e2e_loop(args.steps)
    s1.record(main)
    barrier()
    e2e_ms = max_over_ranks(s0.elapsed_time(s1)) / args.steps
    e2e_value = world * BATCH / (e2e_ms * 1e-3)
    h2d = world * (x_host[0].numel() * 4 + im_host[0].numel() * 4)          # whole job, like `value`
    d2h = world * (out_host[0].numel() * 4 + cnt_host[0].numel() * 4)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel family (tcgen05 implicit-GEMM conv), rank 0 ----------
    peaks, peak_src = measured_peaks()
    ops_t = per_op_times(eng, iters=3)
    conv_ms_eager = sum(t for name, t in ops_t if not is_glue(name))
    total_ms = sum(t for _, t in ops_t)
    conv_launches = sum(1 for name, _ in ops_t if not is_glue(name))
    # the conv launches of one step replayed back to back as their own CUDA graph (same parameters and buffers, glue kernels
    # left out): the kernel family under the conditions of the timed region -- graph replay with programmatic dependent
    # launch -- which per-op events in eager mode cannot give (an event between two launches forbids their overlap)
    conv_graph = eng.capture_steps(lambda name: not is_glue(name))
    for _ in range(3):
        conv_graph.replay()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(main)
    for _ in range(args.steps):
        conv_graph.replay()
    c1.record(main)
    torch.cuda.synchronize()
    conv_ms = c0.elapsed_time(c1) / args.steps
    achieved = eng.conv_flops / (conv_ms * 1e-3) / 1e12
    peak = peaks['bf16_tflops_sustained'] if args.precision == 'bf16' else 75.0
    roofline = {'bound': 'tensor', 'kernel': 'conv_umma_kernel (all %d conv launches of one step)' % conv_launches,
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': ncu_conv_traffic(),
                'traffic_note': 'DRAM read+write bytes of all conv launches of one step (ncu launch list under profiles/); '
                                'algorithmic minimum in min_bytes_per_step',
                'min_bytes_per_step': sum(v['bytes'] for k, v in eng.step_info.items() if not is_glue(k)),
                'peak_source': peak_src + (' bf16_tflops_sustained' if args.precision == 'bf16' else ' (nominal fp32 SIMT)'),
                'flops_per_step': eng.conv_flops, 'kernel_ms_per_step': conv_ms, 'share_of_step': conv_ms / ms_step,
                'timing': 'CUDA events around %d replays of a CUDA graph holding the %d conv launches of one step' % (args.steps, conv_launches),
                'kernel_ms_per_step_eager_events': conv_ms_eager, 'achieved_eager_events': eng.conv_flops / (conv_ms_eager * 1e-3) / 1e12}
    os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(REPO, 'gpurun_out', 'per_op_ms.json'), 'w') as f:
        json.dump({'ops': ops_t, 'total_ms': total_ms, 'graph_ms_per_step': ms_step, 'info': eng.step_info}, f, indent=1)

    # ---- CPU baseline (oracle port), bounded sample -------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, ms, cores = cpu_reference_rate(2, 1)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': '2 forwards of 1 image 608x608 (of the 32-image batch), torch CPU fp32, %d threads' % cores}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': args.precision, 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': world * BATCH, 'l2_policy': 'inputs_exceed_l2 '
                       '(141.9 MB batch, multi-GB activations per step vs 126 MB L2)', 'cuda_graph': True, 'postprocess': eng.postprocess_impl,
                       'nms_candidates_per_image': cand_mean, 'detections_per_image': float(counts.float().mean()),
                       'sharding': 'batch-sharded replicas, no collective'},
            'clocks': clocks, 'gpu_launches': eng.launches_per_run * args.steps * world,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': e2e_ms, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h},
            'roofline': roofline, 'cpu_baseline': cpu, 'matrix_nms': matrix_nms_isolation(dev, world == 1 and not args.no_cpu_baseline),
            'loaded_library': _lib.LIB_PATH}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
