"""Where a movie frame's wall time goes (C5: bench-spheres orbit, 1280x720, 32 passes) and what the overlap buys."""
import sys, time
sys.path.insert(0, ".")
from clive2_b200 import runtime, scene, movie
from clive2_b200.renderer import Renderer

dev = runtime.Device(0)
W, H, S, T = 1280, 720, 32, 120
movie.render_frame(dev, "bench-spheres", W, H, 2, 0, T, seed=1)  # warm up (soup cache, CUDA context)
for f in (1, 2):
    t = [time.time()]
    def lap():
        t.append(time.time()); return t[-1] - t[-2]
    arrays = movie.frame_arrays("bench-spheres", W, H, f, T); a = lap()
    sc = scene.scene_from_arrays(arrays, dev); dev.synchronize() if hasattr(dev, "synchronize") else None; b = lap()
    r = Renderer(sc, seed=1); c = lap()
    for _ in range(S):
        r.run_sample()
    img = r.image; d = lap()
    r.release(); sc.release(); e = lap()
    movie.write_image("/tmp/f.png", img); g = lap()
    print(f"frame {f}: arrays+bvh {a:.3f}  upload+accel {b:.3f}  renderer {c:.3f}  passes+image {d:.3f}  release {e:.3f}  png {g:.3f}", flush=True)
for overlap in (False, True):
    t0 = time.time()
    movie.render_movie(dev, "bench-spheres", W, H, S, range(10, 22), T, seed=1,
                       sink=lambda f, img: movie.write_image(f"/tmp/frame_{f}.png", img), overlap=overlap)
    print(f"overlap={overlap}: {(time.time() - t0) / 12:.3f} s/frame", flush=True)
