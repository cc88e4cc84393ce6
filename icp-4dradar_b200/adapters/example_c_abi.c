/* example_c_abi.c — the scan-to-map loop of the reference (src/radar_odometry.cpp:344-421) written against the C ABI alone:
 * what a host program in any language with a C FFI does. Plain C99, no CUDA headers. Build: make -C icp-4dradar_b200/adapters
 * Run (needs a B200): ./example_c_abi  — registers a few synthetic frames and prints the poses. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "icp4r.h"

static float frand(unsigned* s) { /* xorshift, deterministic */
    *s ^= *s << 13;
    *s ^= *s >> 17;
    *s ^= *s << 5;
    return (float)(*s & 0xFFFFFF) / (float)0x1000000;
}

/* a ground plane and two walls, seen from a sensor at (x0, 0, 1.5), in the sensor frame */
static void make_scan(float* xyzw, int n, float x0, unsigned seed) {
    unsigned s = seed * 2654435761u + 1u;
    for (int i = 0; i < n; ++i) {
        float x = (frand(&s) - 0.5f) * 80.f, y = (frand(&s) - 0.5f) * 80.f, z;
        const int kind = i % 4;
        if (kind < 2) z = 0.f;                                   /* ground */
        else if (kind == 2) { y = 12.f; z = frand(&s) * 4.f; }   /* wall along x */
        else { x = 25.f; z = frand(&s) * 4.f; }                  /* wall along y (makes x observable) */
        xyzw[4 * i + 0] = x - x0;
        xyzw[4 * i + 1] = y;
        xyzw[4 * i + 2] = z - 1.5f;
        xyzw[4 * i + 3] = 1.f;
    }
}

int main(void) {
    icp4r_handle h;
    if (icp4r_create(0, &h) != ICP4R_OK) {
        fprintf(stderr, "icp4r_create: %s\n", icp4r_last_error(NULL));
        return 1;
    }
    const int n = 3000;
    float* scan = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    icp4r_opts o;
    icp4r_default_opts(&o);
    o.residual = ICP4R_P2PLANE_KNN;
    o.k = 5;
    o.max_iterations = 20;
    o.max_corr_dist = 2.0;
    double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 1.5, 0, 0, 0, 1}; /* first pose: sensor 1.5 m above the ground */
    icp4r_map_build(h, NULL, 0, ICP4R_HOST, 0.f);
    for (int f = 0; f < 10; ++f) {
        make_scan(scan, n, 0.4f * (float)f, (unsigned)f);
        icp4r_result r;
        /* register against the map from the previous pose, then insert the scan at the estimated pose */
        if (icp4r_odometry_step(h, scan, n, ICP4R_HOST, &o, 0, T, &r) != ICP4R_OK) {
            fprintf(stderr, "frame %d: %s\n", f, icp4r_last_error(h));
            return 1;
        }
        int32_t size = 0, valid = 0;
        icp4r_map_size(h, &size, &valid);
        printf("frame %d: x = %.3f m (truth %.3f), %d correspondences, map %d points\n", f, T[3], 0.4 * f, r.n_corr, size);
    }
    free(scan);
    icp4r_destroy(h);
    return 0;
}
