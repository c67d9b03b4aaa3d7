import os, sys, time, cProfile, pstats, io
import torch
sys.path.insert(0, "/root/repo")
from segmantic_b200.seg.monai_unet import Net, predict_volume
from segmantic_b200.synthetic import synthetic_state_dict, synthetic_volume
dev = torch.device("cuda:0")
sd = synthetic_state_dict(3, 1, 10, seed=0)
net = Net(num_classes=10, num_channels=1, spatial_dims=3); net.load_state_dict(sd); net.to(dev)
host = synthetic_volume((256, 256, 256), seed=1).contiguous().pin_memory()
kw = dict(overlap=0.5, mode="gaussian", sw_batch_size=32, precision="bf16", crop_foreground=False)
ts = []
for i in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lab = predict_volume(net, host, None, (), **kw)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("per-call ms:", [round(t, 2) for t in ts])
pr = cProfile.Profile(); pr.enable()
for i in range(3): lab = predict_volume(net, host, None, (), **kw)
pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
