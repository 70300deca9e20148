"""First-contact probe for the GPU box: runs each kernel family once, prints max errors (never asserts)."""
import sys, os, time, traceback
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'pytorch-ppyolo_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from ppyolo_b200 import ops, synth, _lib
from ppyolo_b200._lib import PPY_F32, PPY_BF16
from oracle import ppyolo_ref as ref
print(torch.cuda.get_device_name(0), 'bf16 path supported:', _lib.lib.ppy_conv_bf16_supported(), flush=True)
DEV = 'cuda'
def rnd(t): return t.to(torch.bfloat16).float()
def conv(code, n, cin, cout, k, stride, hw, out_f32=True):
    g = torch.Generator().manual_seed(1)
    x = rnd(torch.randn((n, cin, hw, hw), generator=g)); w = rnd(torch.randn((cout, cin, k, k), generator=g) / (cin*k*k)**0.5)
    want = torch.nn.functional.conv2d(x, w, None, stride, (k-1)//2)
    packed = ops.pack_weight(w.to(DEV), code)
    xh = ops.to_nhwc(x.to(DEV), code, packed[1])
    y = ops.conv_nhwc(xh, packed, cin, cout, k, stride, (k-1)//2, torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV), 0, code, out_code=PPY_F32)
    torch.cuda.synchronize()
    got = ops.from_nhwc(y, cout).cpu()
    err = (got - want).abs().max().item()
    print('conv code=%d n=%d cin=%d cout=%d k=%d s=%d hw=%d  max_err=%.3e  scale=%.3e' % (code, n, cin, cout, k, stride, hw, err, want.abs().max().item()), flush=True)
    if err > 1e-2 * want.abs().max().item():
        print(' got ', got[0, :4, 0, :4].numpy().round(3).tolist()); print(' want', want[0, :4, 0, :4].numpy().round(3).tolist())
        d = (got - want).abs()
        print(' err by channel block:', [round(d[:, i:i+8].max().item(), 3) for i in range(0, min(cout, 64), 8)])
        print(' err by pixel row:', [round(d[0, :, i].max().item(), 3) for i in range(min(hw // stride, 8))])
for args in [(PPY_F32, 1, 64, 64, 1, 1, 16), (PPY_F32, 2, 8, 32, 3, 2, 12), (PPY_BF16, 1, 64, 64, 1, 1, 16), (PPY_BF16, 1, 64, 32, 1, 1, 16),
             (PPY_BF16, 1, 128, 128, 1, 1, 16), (PPY_BF16, 2, 64, 256, 3, 1, 16), (PPY_BF16, 1, 256, 512, 3, 2, 20)]:
    try: conv(*args)
    except Exception: traceback.print_exc()
try:
    b, s = synth.nms_inputs(10000, 80, seed=0)
    got = ops.matrix_nms_batched(b[None].to(DEV), s[None].to(DEV), 0.01, 0.01, 500, 100)[0].cpu().numpy()
    want = ref.matrix_nms(b.numpy(), s.numpy(), 0.01, 0.01, 500, 100)
    print('nms shapes', got.shape, want.shape, 'label mismatches', int((got[:, 0] != want[:, 0]).sum()) if got.shape == want.shape else -1,
          'max score err', float(np.abs(got[:, 1] - want[:, 1]).max()) if got.shape == want.shape else -1, flush=True)
    # timing C5 at batch 1 and 32
    for bs in (1, 32):
        B = b[None].repeat(bs, 1, 1).to(DEV).contiguous(); S = s[None].repeat(bs, 1, 1).to(DEV).contiguous()
        out = torch.empty((bs, 100, 6), device=DEV); cnt = torch.empty(bs, dtype=torch.int32, device=DEV); ws = ops.nms_workspace(bs, 10000, 80, B.device)
        for _ in range(3): ops.matrix_nms_launch(B, S, out, cnt, ws, 0.01, 0.01, 500, 100, False, 2.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.matrix_nms_launch(B, S, out, cnt, ws, 0.01, 0.01, 500, 100, False, 2.0)
        e1.record(); torch.cuda.synchronize()
        print('nms C5 bs=%d: %.1f us/batch, %.2f us/img' % (bs, e0.elapsed_time(e1) * 1000 / 20, e0.elapsed_time(e1) * 1000 / 20 / bs), flush=True)
except Exception: traceback.print_exc()
# conv throughput probe
def bench_conv(code, n, cin, cout, k, stride, hw, iters=20):
    x = torch.randn((n, hw, hw, cin), device=DEV).to(ops.torch_dtype(code))
    w = torch.randn((cout, cin, k, k), device=DEV) * 0.05
    packed = ops.pack_weight(w, code)
    sc, sh = torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV)
    out = None
    for _ in range(3): out = ops.conv_nhwc(x, packed, cin, cout, k, stride, (k-1)//2, sc, sh, 1, code, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.conv_nhwc(x, packed, cin, cout, k, stride, (k-1)//2, sc, sh, 1, code, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ho = (hw + 2*((k-1)//2) - k)//stride + 1
    fl = 2.0 * n * ho * ho * cout * cin * k * k
    print('bench code=%d n=%d cin=%d cout=%d k=%d s=%d hw=%d: %.3f ms  %.1f TFLOP/s' % (code, n, cin, cout, k, stride, hw, ms, fl / ms / 1e9), flush=True)
for args in [(PPY_BF16, 32, 256, 256, 3, 1, 38), (PPY_BF16, 32, 1024, 256, 1, 1, 38), (PPY_BF16, 32, 256, 1024, 1, 1, 38), (PPY_BF16, 32, 64, 64, 3, 1, 152),
             (PPY_BF16, 32, 64, 256, 1, 1, 152), (PPY_BF16, 32, 512, 1024, 3, 1, 19), (PPY_BF16, 32, 128, 128, 3, 1, 76), (PPY_F32, 8, 256, 256, 3, 1, 38)]:
    try: bench_conv(*args)
    except Exception: traceback.print_exc()
