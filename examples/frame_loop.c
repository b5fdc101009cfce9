/* frame_loop.c — the per-frame loop of the reference's viewer (main.cpp: write the stimulation, Animation(), read the
 * positions) against the C-ABI, with the asynchronous I/O calls so that the copies of frame k overlap the step of frame k+1.
 *
 *   gcc -O2 -I include examples/frame_loop.c -L sph_sm_monodomain_b200 -lsphsm_b200 -Wl,-rpath,$PWD/sph_sm_monodomain_b200 -o frame_loop
 *   ./frame_loop [nx ny nz frames]
 *
 * Without a CUDA device sphsm_create fails (there is no CPU fallback) and the program says so and exits 1. */
#include <stdio.h>
#include <stdlib.h>

#include "sphsm_b200.h"

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != SPHSM_OK) {                                                            \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, sphsm_last_error(h));     \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

int main(int argc, char **argv) {
    const int nx = argc > 3 ? atoi(argv[1]) : 40, ny = argc > 3 ? atoi(argv[2]) : 20, nz = argc > 3 ? atoi(argv[3]) : 20;
    const int frames = argc > 4 ? atoi(argv[4]) : 100;
    const int n = nx * ny * nz;
    const float s = 0.04f * 0.9f; /* the lattice spacing of main.cpp's init_cube */
    sphsm_handle *h = NULL;
    sphsm_params p;
    sphsm_default_params(&p);
    p.capacity = n;
    p.world[0] = (nx + 2) * s + 0.08f;
    p.world[1] = (ny + 2) * s + 0.08f;
    p.world[2] = (nz + 2) * s + 0.08f;
    p.diagnostics = 0;
    CHECK(sphsm_create(&p, &h));

    float *xyz = (float *)malloc((size_t)n * 3 * sizeof(float));
    float *stim = (float *)calloc((size_t)n, sizeof(float));
    float *out[2] = {(float *)malloc((size_t)n * 3 * sizeof(float)), (float *)malloc((size_t)n * 3 * sizeof(float))};
    int i = 0;
    for (int x = 0; x < nx; x++)
        for (int y = 0; y < ny; y++)
            for (int z = 0; z < nz; z++, i++) {
                xyz[3 * i] = 0.04f + x * s;
                xyz[3 * i + 1] = y * s;
                xyz[3 * i + 2] = 0.04f + z * s;
                stim[i] = x < 4 ? 300.0f : 0.0f; /* pace one end */
            }
    CHECK(sphsm_init_fluid(h, xyz, n));
    for (int k = 0; k < frames; k++) {
        if (k == frames / 2)
            for (i = 0; i < n; i++) stim[i] = 0.0f;            /* stimulus off half-way, like main.cpp:329 */
        CHECK(sphsm_set_masks_async(h, NULL, stim, n));        /* H2D of this frame's input */
        CHECK(sphsm_step(h, 1));                               /* Animation() */
        CHECK(sphsm_download_positions_async(h, out[k & 1], n)); /* D2H overlaps the next step */
        if (k) {
            /* frame k-1 is complete in out[(k-1)&1] once the I/O streams have drained up to it; a viewer would draw it here */
        }
    }
    CHECK(sphsm_sync(h));
    double cx = 0.0;
    for (i = 0; i < n; i++) cx += out[(frames - 1) & 1][3 * i];
    printf("frame_loop OK: %d particles, %d frames, mean x = %.6f, %d steps taken\n", n, frames, cx / n, sphsm_total_time_steps(h));
    sphsm_destroy(h);
    free(xyz); free(stim); free(out[0]); free(out[1]);
    return 0;
}
