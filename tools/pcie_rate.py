import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, rep=5):
    fn(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(rep): fn()
    torch.cuda.synchronize()
    return (time.time() - t0) / rep
a = t(lambda: d.copy_(h, non_blocking=True))
b = t(lambda: h.copy_(d, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print("H2D GB/s %.1f  D2H GB/s %.1f  both directions at once: %.1f GB/s each" % (n / a / 1e9, n / b / 1e9, n / c / 1e9))
