"""GPU box helper: a short run of one workload's Lloyd loop (used under ncu).  Not part of the product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cniic_b200 as cb
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind, w, h, k, blobs = {"c2": (cb.POINTS_RGB, 4096, 4096, 256, 192), "c3": (cb.POINTS_XYRGB, 7680, 4320, 2048, 2048),
                        "c1": (cb.POINTS_RGB, 512, 512, 16, 24), "c4": (cb.POINTS_RGB, 1024, 1024, 64, 16)}[wl]
ctx = cb.Context(0)
d = ctx.device_alloc(w * h * 3)
cb.synth_image_device(ctx, d, w, h, 0xC0FFEE + int(wl[1]), blobs)
s = cb.KMeansSession(ctx, kind, k, d, w * h, w=w, h_local=h, on_device=True, flags=int(os.environ.get('KM_FLAGS', '0')))
s.reset()
st = s.run(iters)
print(wl, st.iterations, "iters", st.device_ms, "ms", w * h * st.iterations / st.device_ms / 1e3, "Mpx.iter/s; assign avg", st.assign_ms_avg, "ms")
