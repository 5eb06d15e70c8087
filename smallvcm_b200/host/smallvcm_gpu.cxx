// smallvcm_gpu.cxx — builds the reference's OWN driver (smallvcm.cxx: main, render,
// FullReport, the CLI) with the five VertexCM algorithms served by the GPU renderer, without
// editing a single reference file:
//
//   1. the reference headers are included in smallvcm.cxx's order (src/smallvcm.cxx:25-39);
//   2. VertexCMGpu is defined (header-only, needs AbstractRenderer / Scene / Framebuffer);
//   3. config.hxx — whose factory CreateRenderer names `VertexCM` five times
//      (src/config.hxx:124-138) — is included with `VertexCM` macro-renamed to VertexCMGpu;
//   4. smallvcm.cxx itself is included verbatim; all its header includes are now no-ops
//      through the include guards;
//   5. while it is included, `clock` names a WALL clock: render() budgets `-t <sec>` and reports
//      its time with clock() (src/smallvcm.cxx:74,82,111,150), i.e. process CPU time summed over
//      the OpenMP threads, which says nothing about work done on the GPU while the host threads
//      wait. With the substitution `-t 5` means five seconds of rendering
//      and "done in" / the --report table show elapsed time.
//
// el / pt run the reference's CPU classes (EyeLight, PathTracer) unless SVCM_GPU_PT=1 is set, so
// BASELINE config #1 (`-a pt`, no GPU needed) works from the same binary.
//
// Build (see Makefile next to this file):
//   g++ -O3 -std=c++0x -fopenmp -fno-access-control -I<SmallVCM>/src -I<repo>/include \
//       smallvcm_gpu.cxx -L<repo>/smallvcm_b200/lib -lsvcm
#include <vector>
#include <cmath>
#include <time.h>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <map>
#include <string>

#include "math.hxx"
#include "ray.hxx"
#include "geometry.hxx"
#include "camera.hxx"
#include "framebuffer.hxx"
#include "scene.hxx"
#include "eyelight.hxx"
#include "pathtracer.hxx"
#include "bsdf.hxx"
#include "vertexcm.hxx"
#include "html_writer.hxx"

#include "vertexcm_gpu.hxx"

#define VertexCM VertexCMGpu
#define PathTracer PathTracerGpu     // GPU only with SVCM_GPU_PT=1, else the reference's CPU class
#define EyeLight EyeLightGpu
#include "config.hxx"
#undef VertexCM
#undef PathTracer
#undef EyeLight

namespace svcm_host {
inline clock_t WallClock()
{
    static timespec start = { 0, 0 };
    timespec now;
    clock_gettime(CLOCK_MONOTONIC, &now);
#pragma omp critical(svcm_wall_clock)
    if(start.tv_sec == 0 && start.tv_nsec == 0) start = now;
    const double seconds = double(now.tv_sec - start.tv_sec) + 1e-9 * double(now.tv_nsec - start.tv_nsec);
    return clock_t(seconds * CLOCKS_PER_SEC);
}
}

#define clock svcm_host::WallClock
#include "smallvcm.cxx"
#undef clock
