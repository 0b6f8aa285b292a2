import torch, time, numpy as np
n = 1250 * 1000 * 1000
dev = torch.empty(n, dtype=torch.float32, device='cuda').fill_(1.0)
pin = torch.empty(n, dtype=torch.float32).pin_memory()
cs = torch.cuda.Stream(); hs = torch.cuda.Stream()
x = np.zeros((1250, 500)); xp = torch.from_numpy(x).pin_memory()
def small_h2d(label, src):
    t0 = time.perf_counter()
    with torch.cuda.stream(hs):
        y = src.to('cuda', non_blocking=True)
    hs.synchronize()
    print('  %s small H2D done after %.2f ms' % (label, 1e3 * (time.perf_counter() - t0)))
for slices in (1, 16, 64):
    torch.cuda.synchronize()
    with torch.cuda.stream(cs):
        step = n // slices
        for i in range(slices):
            pin[i * step:(i + 1) * step].copy_(dev[i * step:(i + 1) * step], non_blocking=True)
    time.sleep(0.01)
    print('slices', slices)
    small_h2d('pageable', torch.from_numpy(x))
    small_h2d('pinned', xp)
    t0 = time.perf_counter(); cs.synchronize(); print('  rest of D2H %.1f ms' % (1e3 * (time.perf_counter() - t0)))
