// TEST INFRASTRUCTURE ONLY.  Thin driver linked with the reference's own src/render.cpp (compiled
// against oracle/shim) into oracle/_ref/libref_<tag>.so, so tests and the bench's cpu_baseline leg
// can call the reference kernel in-process instead of through files.  It does what the reference's
// src/main.cpp:37 does -- ICPU_RUN_KF(render, 8, rays, spheres, colors) -- and nothing else.
#include "tikicpulib.h"
#include "common.h" // the generated one in the build directory: WIDTH / HEIGHT / SAMPLES of this build

extern "C" __global__ __aicore__ void render(GM_ADDR rays, GM_ADDR spheres, GM_ADDR colors);

extern "C" {
// Problem size baked into this build (the reference's sizes are compile-time, src/common.h:4-6).
void ref_dims(int32_t *w, int32_t *h, int32_t *s) {
    *w = WIDTH;
    *h = HEIGHT;
    *s = SAMPLES;
}
// rays: float32 SoA [6][N]; spheres: 512 B; colors: float32 SoA [3][N]; N = W*H*S*4.
// nthreads <= 0 -> honour PT_REF_THREADS (default 1).
void ref_render(uint8_t *rays, uint8_t *spheres, uint8_t *colors, int32_t nthreads) {
    if (nthreads > 0) {
        char buf[16];
        std::snprintf(buf, sizeof buf, "%d", nthreads);
        setenv("PT_REF_THREADS", buf, 1);
    }
    const uint32_t blockDim = 8; // src/main.cpp:18
    ICPU_RUN_KF(render, blockDim, rays, spheres, colors);
}
}
