"""Times the three ways of fetching the picture after an iteration (1024x1024 unless given):
float accumulated image (12 B/pixel, what GetFramebuffer needs) vs the device-converted BMP
(3 B/pixel) and RGBE (4 B/pixel) payloads. Host wall time of the blocking C-ABI call."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import smallvcm_b200 as sv  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
r = sv.VertexCM(sv.Scene.cornell_box(1, res, res), "vcm"); r.max_path_length = 10
r.run_iterations(0, 3); r.synchronize()
buf = np.empty((res, res, 3), np.float32)
lib = sv.load_library()
lib.svcm_host_register(buf.ctypes.data, buf.nbytes)
for name, fn in (("float accum (pinned)", lambda: r.accum(buf)), ("float accum (pageable)", lambda: r.accum()),
                 ("bmp payload", lambda: r.image8("bmp")), ("rgbe payload", lambda: r.image8("hdr"))):
    fn()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    print("%-24s %.3f ms" % (name, (time.perf_counter() - t0) / 20 * 1e3))
