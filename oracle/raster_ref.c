/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the neural_renderer forward rasterizer.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  The product path (impersonator_b200/) never links or calls it.
 *
 * A restatement in plain C of
 *   thirdparty/neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu
 *     :40-84   forward_face_index_map_cuda_kernel_1   (per-face inverse matrices)
 *     :86-186  forward_face_index_map_cuda_kernel_2   (per-pixel brute-force z-buffer)
 * with the fp32/FMA shape that nvcc 12.9 (-fmad=true, default flags) gives those kernels
 * on sm_100a.  The contraction pattern below was read off the SASS of the reference file
 * compiled unmodified (oracle/build_ref.sh -> oracle/_ref/libnmr_ref.so), see DESIGN.md:
 *   p      = ((v*is + is) [one FFMA]  - 1) * 0.5
 *   a*b-c*d= fmaf(a, b, -(c*d))
 *   det    = fmaf(p10, p21-p01, fmaf(p20, p01-p11, p00*(p11-p21)))
 *   w      = (fmaf(inv0, xi, inv1*yi)) + inv2
 *   zp     = 1 / ((w0/z0 + w1/z1) + w2/z2)        (IEEE, correctly rounded)
 * Build with -ffp-contract=off so that ONLY the explicit fmaf() calls fuse.
 *
 * Pinned by: tests/golden/teapot.npz (silhouette exact, depth atol 1e-2 -- the reference's own
 * tests/test_rasterize_silhouettes.py:16-35 and tests/test_rasterize_depth.py:37-54) on CPU,
 * and bit-for-bit against oracle/_ref/libnmr_ref.so on the GPU (tests/test_raster_gpu.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

/* kernel_1: rasterize_cuda_kernel.cu:48-83.  faces_inv must be zero-filled by the caller
 * (rasterize.py:165 passes torch.zeros_like(faces)); culled faces leave their row untouched. */
static void face_setup(const float* face, float* inv, int is)
{
    /* :57  return if backside */
    if ((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]))
        return;
    const float fis = (float)is;
    float p[3][2];
    for (int n = 0; n < 3; n++)
        for (int d = 0; d < 2; d++)
            p[n][d] = (fmaf(face[3 * n + d], fis, fis) + -1.0f) * 0.5f;     /* :64 */
    float a[9];
    a[0] = p[1][1] - p[2][1];
    a[1] = p[2][0] - p[1][0];
    a[2] = fmaf(p[1][0], p[2][1], -(p[2][0] * p[1][1]));
    a[3] = p[2][1] - p[0][1];
    a[4] = p[0][0] - p[2][0];
    a[5] = fmaf(p[2][0], p[0][1], -(p[0][0] * p[2][1]));
    a[6] = p[0][1] - p[1][1];
    a[7] = p[1][0] - p[0][0];
    a[8] = fmaf(p[0][0], p[1][1], -(p[1][0] * p[0][1]));
    const float det = fmaf(p[1][0], p[2][1] - p[0][1],
                      fmaf(p[2][0], p[0][1] - p[1][1], p[0][0] * (p[1][1] - p[2][1])));   /* :73-76 */
    for (int k = 0; k < 9; k++) inv[k] = a[k] / det;                                         /* :77-83 */
}

static inline float clamp01(float w)
{
    /* :146  min(max(w, 0.), 1.) evaluated in double; fmax/fmin return the non-NaN operand. */
    double d = fmax((double)w, 0.0);
    d = fmin(d, 1.0);
    return (float)d;
}

/* one pixel of kernel_2 (:103-185) */
static void pixel_eval(const float* faces, const float* faces_inv, int F, int is, float near_, float far_,
                       long i, int32_t* fim, float* wim, float* depth)
{
    const int bn = (int)(i / ((long)is * is));
    const int pn = (int)(i % ((long)is * is));
    const int yi = pn / is, xi = pn % is;
    const float yp = (float)((2. * yi + 1 - is) / is);      /* :113 */
    const float xp = (float)((2. * xi + 1 - is) / is);      /* :114 */
    const float fx = (float)xi, fy = (float)yi;
    const float* face = faces + (size_t)bn * F * 9;
    const float* inv = faces_inv + (size_t)bn * F * 9;
    float depth_min = far_;
    int face_min = -1;
    float wmin[3] = {0, 0, 0};
    for (int fn = 0; fn < F; fn++, face += 9, inv += 9) {
        if ((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]))
            continue;                                       /* :128 */
        if (((yp - face[1]) * (face[3] - face[0]) < (xp - face[0]) * (face[4] - face[1])) ||
            ((yp - face[4]) * (face[6] - face[3]) < (xp - face[3]) * (face[7] - face[4])) ||
            ((yp - face[7]) * (face[0] - face[6]) < (xp - face[6]) * (face[1] - face[7])))
            continue;                                       /* :132-135 */
        float w[3];
        for (int k = 0; k < 3; k++)
            w[k] = fmaf(inv[3 * k + 0], fx, inv[3 * k + 1] * fy) + inv[3 * k + 2];   /* :139-141 */
        float ws = 0.0f;
        for (int k = 0; k < 3; k++) { w[k] = clamp01(w[k]); ws += w[k]; }           /* :144-148 */
        for (int k = 0; k < 3; k++) w[k] /= ws;                                      /* :149-151 */
        const float zp = 1.0f / ((w[0] / face[2] + w[1] / face[5]) + w[2] / face[8]);   /* :153 */
        if (zp <= near_ || far_ <= zp) continue;            /* :154 */
        if (zp < depth_min) {                               /* :159 strict: lowest index wins ties */
            depth_min = zp; face_min = fn;
            wmin[0] = w[0]; wmin[1] = w[1]; wmin[2] = w[2];
        }
    }
    if (face_min >= 0) {                                    /* :174-179 */
        depth[i] = depth_min;
        fim[i] = face_min;
        wim[3 * i + 0] = wmin[0]; wim[3 * i + 1] = wmin[1]; wim[3 * i + 2] = wmin[2];
    }
}

typedef struct {
    const float* faces; float* faces_inv; int B, F, is; float near_, far_;
    int32_t* fim; float* wim; float* depth; int tid, nthreads; int phase;
} job_t;

static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    if (j->phase == 0) {
        const long n = (long)j->B * j->F;
        for (long i = j->tid; i < n; i += j->nthreads) face_setup(j->faces + i * 9, j->faces_inv + i * 9, j->is);
    } else {
        /* interleave rows across threads so covered rows are shared evenly */
        const long rows = (long)j->B * j->is;
        for (long r = j->tid; r < rows; r += j->nthreads)
            for (long i = r * j->is; i < (r + 1) * j->is; i++)
                pixel_eval(j->faces, j->faces_inv, j->F, j->is, j->near_, j->far_, i, j->fim, j->wim, j->depth);
    }
    return NULL;
}

static int g_threads = 0;

int lwb_oracle_num_threads(void)
{
    if (g_threads <= 0) {
        const char* e = getenv("LWB_ORACLE_THREADS");
        long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
        if (n < 1) n = 1;
        if (n > 256) n = 256;
        g_threads = (int)n;
    }
    return g_threads;
}

void lwb_oracle_set_num_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }

/*
 * faces      f32 [B,F,3,3]  NDC, +y up (as passed to rasterize_cuda.forward_face_index_map)
 * faces_inv  f32 [B,F,3,3]  out (zero-filled here, then kernel_1)
 * fim        i32 [B,is,is]  pre-filled by caller (-1); written only where covered   (:174-185)
 * wim        f32 [B,is,is,3] pre-filled by caller (0)
 * depth      f32 [B,is,is]  pre-filled by caller (far)
 * Row order is the kernel's (row 0 = bottom); the flip of rasterize.py:334-338 is the caller's.
 */
void lwb_oracle_forward_face_index_map(const float* faces, int B, int F, int is, float near_, float far_,
                                       int32_t* fim, float* wim, float* depth, float* faces_inv)
{
    memset(faces_inv, 0, sizeof(float) * (size_t)B * F * 9);
    const int nt = lwb_oracle_num_threads();
    pthread_t th[256];
    job_t jobs[256];
    for (int phase = 0; phase < 2; phase++) {
        for (int t = 0; t < nt; t++) {
            job_t j = {faces, faces_inv, B, F, is, near_, far_, fim, wim, depth, t, nt, phase};
            jobs[t] = j;
            if (nt > 1) pthread_create(&th[t], NULL, worker, &jobs[t]);
            else worker(&jobs[t]);
        }
        if (nt > 1) for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    }
}
