import os, time, torch
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
dev = torch.device("cuda:0")
a = torch.zeros(1, device=dev)
sA, sB = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
a.add_(1.0)                              # load the kernel before the experiment
torch.cuda.synchronize()
eB = torch.cuda.Event()
t0 = time.time()
with torch.cuda.stream(sA):
    torch.cuda._sleep(int(2e9))          # ~1 s spin on stream A
t1 = time.time()
with torch.cuda.stream(sB):
    a.add_(1.0)
    eB.record(sB)
t2 = time.time()
eB.synchronize()
print(f"host: sleep enqueued in {t1 - t0:.4f}s, add enqueued in {t2 - t1:.4f}s")
print(f"stream B finished at +{time.time() - t0:.3f}s (concurrent if << 1 s)")
torch.cuda.synchronize()
print(f"all finished at +{time.time() - t0:.3f}s")
