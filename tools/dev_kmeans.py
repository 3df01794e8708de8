"""Developer script (GPU box): quick parity + timing of the K-means kernels. Not part of the product or tests."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cniic_b200 as cb
import oracle as O

ctx = cb.Context()
ok = True
for (w, h, k, tie) in [(64, 48, 16, 0), (200, 150, 64, 1), (512, 512, 16, 0), (333, 257, 256, 0)]:
    img = cb.synth_image_host(w, h, 0xC0FFEE + k, 24)
    for mi in (1, 3, 0):
        g = ctx.kmeans_rgb(img.reshape(-1, 3), k, max_iters=mi, tie=tie)
        o = O.kmeans_rgb(img.reshape(-1, 3), k, mode=O.MODE_EXACT, tie=tie, max_iters=mi)
        same = np.array_equal(g.centroids, o.centroids) and np.array_equal(g.assign, o.assign) and g.iterations == o.iterations
        print("rgb", w, h, k, tie, mi, "iters", g.iterations, o.iterations, "OK" if same else "MISMATCH", flush=True)
        ok &= same
        g = ctx.kmeans_xyrgb(img, k, max_iters=mi, tie=tie)
        o = O.kmeans_xyrgb(img, k, mode=O.MODE_EXACT, tie=tie, max_iters=mi)
        same = np.array_equal(g.centroids, o.centroids) and np.array_equal(g.assign, o.assign) and g.iterations == o.iterations
        if not same:
            print("  cen diff", np.abs(g.centroids - o.centroids).max(), "asg diff", int((g.assign != o.assign).sum()))
        print("xyrgb", w, h, k, tie, mi, "iters", g.iterations, o.iterations, "OK" if same else "MISMATCH", flush=True)
        ok &= same
print("PARITY", "GREEN" if ok else "RED")

def timed(kind, w, h, k, iters=10, nblobs=192):
    n = w * h
    d = ctx.device_alloc(n * 3)
    cb.synth_image_device(ctx, d, w, h, 0xC0FFEE, nblobs)
    ctx.sync()
    s = cb.KMeansSession(ctx, kind, k, d, n, w=w, h_local=h, on_device=True)
    for rep in range(3):
        s.reset()
        st = s.run(iters)
        print(f"kind={kind} {w}x{h} k={k}: {st.iterations} iters {st.device_ms:.3f} ms -> {n*st.iterations/st.device_ms/1e3:.1f} Mpx.iter/s  launches={st.gpu_launches} moved_last={st.moved_last}", flush=True)
    s.close()
    ctx.device_free(d)

timed(cb.POINTS_RGB, 512, 512, 16)
timed(cb.POINTS_RGB, 4096, 4096, 256)
timed(cb.POINTS_RGB, 4096, 4096, 64)
timed(cb.POINTS_XYRGB, 7680, 4320, 2048, iters=4, nblobs=2048)
timed(cb.POINTS_XYRGB, 4096, 4096, 256, iters=4)
