import time, torch
n_h2d, n_d2h = 327_696_384, 82_057_216
a = torch.empty(n_h2d, dtype=torch.uint8).pin_memory(); b = torch.empty(n_d2h, dtype=torch.uint8).pin_memory()
da = torch.empty(n_h2d, dtype=torch.uint8, device="cuda"); db = torch.empty(n_d2h, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): b.copy_(db, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for _ in range(2): run(True, True, 2)
t = run(True, False); print("H2D alone  %.2f ms  %.1f GB/s" % (1e3 * t, n_h2d / t / 1e9))
t = run(False, True); print("D2H alone  %.2f ms  %.1f GB/s" % (1e3 * t, n_d2h / t / 1e9))
t = run(True, True); print("both       %.2f ms  (H2D-equivalent %.1f GB/s)" % (1e3 * t, n_h2d / t / 1e9))
