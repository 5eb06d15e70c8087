// smallvcm_gpu.cxx — builds the reference's OWN driver (smallvcm.cxx: main, render,
// FullReport, the CLI) with the five VertexCM algorithms served by the GPU renderer, without
// editing a single reference file:
//
//   1. the reference headers are included in smallvcm.cxx's order (src/smallvcm.cxx:25-39);
//   2. VertexCMGpu is defined (header-only, needs AbstractRenderer / Scene / Framebuffer);
//   3. config.hxx — whose factory CreateRenderer names `VertexCM` five times
//      (src/config.hxx:124-138) — is included with `VertexCM` macro-renamed to VertexCMGpu;
//   4. smallvcm.cxx itself is included verbatim; all its header includes are now no-ops
//      through the include guards.
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

#include "smallvcm.cxx"
