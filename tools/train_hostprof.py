"""Host-side profile (cProfile) of Trainer.step at the C4 shape: which call blocks the host."""
import cProfile, pstats, io, os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'pytorch-ppyolo_b200')); sys.path.insert(0, REPO)
import torch
import bench
from ppyolo_b200 import synth, targets as tg
from ppyolo_b200.trainer import Trainer
dev = torch.device('cuda', 0)
model, cfg = bench.build_model(bench.ARCH, train=True)
model = model.to(dev); model.train_precision = 'bf16'
trainer = Trainer(model, cfg, graph=True)
x = synth.images(8, 608, seed=20).to(dev)
gb, gc, gs = tg.synthetic_ground_truth(8, seed=30)
targets = [torch.from_numpy(t).to(dev) for t in tg.gt2yolo_target(gb, gc, gs, h=608, w=608, **cfg.gt2YoloTarget)]
gb, gc, gs = (torch.from_numpy(v).to(dev) for v in (gb, gc, gs))
for _ in range(4):
    trainer.step(x, gb, gc, gs, targets)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    trainer.step(x, gb, gc, gs, targets)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(14); print(s.getvalue()[:4000])
