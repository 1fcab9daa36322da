/* Plain-C caller of the C ABI (include/mamri_b200.h): what a non-Python host would write behind
 * MamriLogic.volume_threshold_segmentation (Mamri/Mamri.py:1304-1323).  Reads a raw uint16 volume
 * [nz][ny][nx] from a file, runs the detection from/to host buffers, prints the markers the way the reference
 * labels them (Mamri.py:1317) and the body label.
 *
 *   gcc -std=c99 -I include examples/detect_host.c -o detect_host -L mamri_pose_estimation_b200 -lmamri_b200 \
 *       -Wl,-rpath,$PWD/mamri_pose_estimation_b200
 *   ./detect_host volume.u16 nx ny nz sx sy sz
 */
#include <stdio.h>
#include <stdlib.h>

#include "mamri_b200.h"

int main(int argc, char** argv) {
    if (argc != 8) { fprintf(stderr, "usage: %s volume.u16 nx ny nz sx sy sz\n", argv[0]); return 2; }
    const int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]);
    const size_t n = (size_t)nx * ny * nz;
    uint16_t* vol = (uint16_t*)malloc(n * sizeof(uint16_t));
    uint8_t* body = (uint8_t*)malloc(n);
    FILE* f = fopen(argv[1], "rb");
    if (!vol || !body || !f || fread(vol, sizeof(uint16_t), n, f) != n) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);

    mamri_ctx* ctx = NULL;
    int rc = mamri_create(&ctx, 0, nx, ny, nz, 0, 0);
    if (rc != MAMRI_OK) { fprintf(stderr, "mamri_create: %d %s\n", rc, mamri_last_error(NULL)); return 1; }
    mamri_volume_desc d = {nx, ny, nz, MAMRI_U16, {atof(argv[5]), atof(argv[6]), atof(argv[7])}, {0, 0, 0}, {1, 0, 0, 0, 1, 0, 0, 0, 1}};
    mamri_params p;
    mamri_default_params(&p);                 /* 65.0 .. 65535, ball radius 2, 6-connected, 50 .. 1500 mm^3 */
    rc = mamri_detect_host_async(ctx, &d, vol, &p, body, NULL);
    mamri_summary s;
    static mamri_marker markers[256];
    if (rc == MAMRI_OK) rc = mamri_detect_collect(ctx, &s, markers, 256);
    if (rc != MAMRI_OK) { fprintf(stderr, "detect: %d %s\n", rc, mamri_last_error(ctx)); mamri_destroy(ctx); return 1; }
    printf("labels %u markers %u body %u body_voxels %llu\n", s.n_labels, s.n_markers, s.body_label, (unsigned long long)s.body_count);
    for (uint32_t i = 0; i < s.n_markers; ++i)
        printf("M_%u_%.0fmm3 ras %.6f %.6f %.6f count %llu\n", markers[i].label, markers[i].volume_mm3, markers[i].centroid_ras[0],
               markers[i].centroid_ras[1], markers[i].centroid_ras[2], (unsigned long long)markers[i].count);
    unsigned long long in_body = 0;
    for (size_t i = 0; i < n; ++i) in_body += body[i];
    printf("body_mask_voxels %llu\n", in_body);
    mamri_destroy(ctx);
    free(vol); free(body);
    return 0;
}
