"""Host<->device copy rates of the box (pinned memory), alone and in both directions at once: the denominator of the
end-to-end inverse_map figure (3 KB per sample cross PCIe)."""
import json
import time

import torch

n = 1 << 30                                               # 1 GiB
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device='cuda')
d_b = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


out = {'bytes': n, 'h2d_gbs': n / timed(h2d) / 1e9, 'd2h_gbs': n / timed(d2h) / 1e9}
t = timed(both)
out['both_s'] = t
out['both_total_gbs'] = 2 * n / t / 1e9
# chunked (168 MB pieces, as the inverse pipeline issues them)
c = 168 * 1024 * 1024


def h2d_chunks():
    with torch.cuda.stream(s1):
        for o in range(0, n - c + 1, c):
            d_a[o:o + c].copy_(h_in[o:o + c], non_blocking=True)


out['h2d_chunked_gbs'] = (n // c) * c / timed(h2d_chunks) / 1e9
print(json.dumps(out))
