import torch, time
n = 1250 * 1000 * 1000
dev = torch.empty(n, dtype=torch.float32, device='cuda').fill_(1.0)
pin = torch.empty(n, dtype=torch.float32).pin_memory()
a = torch.randn(8192, 8192, dtype=torch.float64, device='cuda'); b = torch.randn(8192, 8192, dtype=torch.float64, device='cuda')
cs = torch.cuda.Stream()
def t(fn, label):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); print(label, '%.1f ms' % (1e3 * (time.perf_counter() - t0)))
def d2h():
    with torch.cuda.stream(cs): pin.copy_(dev, non_blocking=True)
def gemm():
    for _ in range(2): torch.matmul(a, b)
t(d2h, 'warm d2h'); t(gemm, 'warm gemm')
t(d2h, 'd2h 5GB alone'); t(gemm, 'gemm alone')
def both(): d2h(); gemm()
t(both, 'd2h + gemm overlapped')
x = torch.empty(1250, 500, dtype=torch.float64).numpy()
def pageable_after_gemm():
    gemm(); t0 = time.perf_counter(); y = torch.from_numpy(x).to('cuda', non_blocking=True); print('  pageable H2D call returned after %.1f ms' % (1e3 * (time.perf_counter() - t0)))
t(pageable_after_gemm, 'gemm then pageable h2d')
