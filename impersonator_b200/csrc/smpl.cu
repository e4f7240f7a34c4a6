// SMPL body model: pose -> vertices (linear blend skinning), the producer of `verts` for every frame
// (SURVEY.md 8f rank 1).  Replaces SMPL.forward (networks/batch_smpl.py:285-375) with its helpers
// batch_rodrigues (:64-101) and batch_global_rigid_transformation (:129-218).
//
// Three launches per batch, all latency/HBM-bound (the whole model is 19 MB, read once per 8 frames):
//   k_smpl_chain  one warp per frame: 24 Rodrigues rotations, pose feature (R - I, 207 values), joints of
//                 the shaped template (J = J0 + JS*beta: J_regressor is linear, so it is folded into the
//                 model at load time), the kinematic chain and the relative transforms A [24][3x4].
//   k_smpl_skin   32 vertices x up to 8 frames per block: shape blend (10 terms) + pose blend (207 terms, the
//                 K range split over 4 thread slices, posedirs read coalesced and shared by the 8 frames),
//                 then T = sum_j w[v][j] A[j] and verts = T [v;1].
//   k_smpl_joints cocoplus joints = joint_regressor^T verts (19 x 3 reductions over 6890 vertices per frame).
#include "common.cuh"

namespace {

constexpr int NJ = 24;            // SMPL kinematic joints
constexpr int NPF = 207;          // pose feature = 23 * 9
constexpr int FB = 8;             // frames per block pass in k_smpl_skin
constexpr int VB = 32;            // vertices per block
constexpr int KS = 4;             // K slices of the pose blend
constexpr int MAX_BETAS = 16;

__global__ void __launch_bounds__(32) k_smpl_chain(
        const float* __restrict__ beta, const float* __restrict__ theta, int num_betas,
        const float* __restrict__ j_template, const float* __restrict__ j_shapedirs,
        const int* __restrict__ parents, int rotate_base,
        float* __restrict__ pose_feature, float* __restrict__ A_out, float* __restrict__ Rs_out,
        float* __restrict__ J_out)
{
    const int b = blockIdx.x, lane = threadIdx.x;
    __shared__ float R[NJ][9];
    __shared__ float J[NJ][3];
    __shared__ float G[NJ][12];          // global transforms (rows 0..2 of the 4x4)
    __shared__ int par[NJ];
    if (lane < NJ) {
        par[lane] = parents[lane];
        // batch_rodrigues (batch_smpl.py:64-101): angle = ||theta + 1e-8||, r = theta / angle
        const float tx = theta[(size_t)b * 72 + lane * 3], ty = theta[(size_t)b * 72 + lane * 3 + 1],
                    tz = theta[(size_t)b * 72 + lane * 3 + 2];
        const float ax = tx + 1e-8f, ay = ty + 1e-8f, az = tz + 1e-8f;
        const float angle = sqrtf(ax * ax + ay * ay + az * az);
        const float rx = tx / angle, ry = ty / angle, rz = tz / angle;
        const float c = cosf(angle), s = sinf(angle), oc = 1.f - c;
        float m[9];
        m[0] = c + oc * (rx * rx);      m[1] = oc * (rx * ry) + s * -rz; m[2] = oc * (rx * rz) + s * ry;
        m[3] = oc * (ry * rx) + s * rz; m[4] = c + oc * (ry * ry);       m[5] = oc * (ry * rz) + s * -rx;
        m[6] = oc * (rz * rx) + s * -ry; m[7] = oc * (rz * ry) + s * rx; m[8] = c + oc * (rz * rz);
#pragma unroll
        for (int k = 0; k < 9; k++) {
            R[lane][k] = m[k];
            if (Rs_out) Rs_out[((size_t)b * NJ + lane) * 9 + k] = m[k];
            if (lane > 0) pose_feature[(size_t)b * NPF + (lane - 1) * 9 + k] = m[k] - ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f);
        }
    }
    // joints of the shaped template: J = J0 + JS * beta
    for (int i = lane; i < NJ * 3; i += 32) {
        float acc = 0.f;
        for (int k = 0; k < num_betas; k++) acc = fmaf(beta[(size_t)b * num_betas + k], j_shapedirs[i * num_betas + k], acc);
        J[i / 3][i % 3] = acc + j_template[i];
    }
    __syncwarp();
    // kinematic chain (batch_smpl.py:180-194): G_0 = [R_0 (x rot_x) | J_0], G_i = G_parent * [R_i | J_i - J_parent]
    const int r = lane >> 2, cc = lane & 3;         // lanes 0..11: element (r, cc) of the 3x4
    if (lane < 12) {
        float v;
        if (cc < 3) v = R[0][r * 3 + cc] * ((rotate_base && cc > 0) ? -1.f : 1.f);
        else v = J[0][r];
        G[0][lane] = v;
    }
    __syncwarp();
    for (int i = 1; i < NJ; i++) {
        const int p = par[i];
        float v = 0.f;
        if (lane < 12) {
            const float g0 = G[p][r * 4], g1 = G[p][r * 4 + 1], g2 = G[p][r * 4 + 2], g3 = G[p][r * 4 + 3];
            if (cc < 3) v = g0 * R[i][cc] + g1 * R[i][3 + cc] + g2 * R[i][6 + cc];
            else v = g0 * (J[i][0] - J[p][0]) + g1 * (J[i][1] - J[p][1]) + g2 * (J[i][2] - J[p][2]) + g3;
        }
        __syncwarp();
        if (lane < 12) G[i][lane] = v;
        __syncwarp();
    }
    // relative transforms (batch_smpl.py:204-216): A = G - [0 | G_rot * J]
    for (int i = lane; i < NJ * 12; i += 32) {
        const int j = i / 12, e = i % 12, rr = e >> 2, c4 = e & 3;
        float v = G[j][e];
        if (c4 == 3) v -= G[j][rr * 4] * J[j][0] + G[j][rr * 4 + 1] * J[j][1] + G[j][rr * 4 + 2] * J[j][2];
        A_out[(size_t)b * NJ * 12 + i] = v;
    }
    if (J_out) for (int i = lane; i < NJ * 3; i += 32) J_out[(size_t)b * NJ * 3 + i] = G[i / 3][(i % 3) * 4 + 3];
}

__global__ void __launch_bounds__(VB * 3 * KS) k_smpl_skin(
        const float* __restrict__ beta, int num_betas, int batch, int V,
        const float* __restrict__ v_template, const float* __restrict__ shapedirs, const float* __restrict__ posedirs,
        const float* __restrict__ weights, const float* __restrict__ pose_feature, const float* __restrict__ A,
        float* __restrict__ verts)
{
    constexpr int NC = VB * 3;                      // coordinates per block
    const int c = threadIdx.x % NC, ks = threadIdx.x / NC;
    const int v0 = blockIdx.x * VB;
    const long gc = (long)v0 * 3 + c;               // global coordinate index (v*3 + d)
    const long VC = (long)V * 3;
    const bool live = gc < VC;
    __shared__ float s_pf[FB][NPF + 1];
    __shared__ float s_beta[FB][MAX_BETAS];
    __shared__ float s_A[FB][NJ * 12];
    __shared__ float s_red[KS][FB][NC];
    for (int f0 = 0; f0 < batch; f0 += FB) {
        const int nf = min(FB, batch - f0);
        __syncthreads();
        for (int i = threadIdx.x; i < FB * NPF; i += blockDim.x) {
            const int f = i / NPF, k = i % NPF;
            s_pf[f][k] = f < nf ? pose_feature[(size_t)(f0 + f) * NPF + k] : 0.f;
        }
        for (int i = threadIdx.x; i < FB * num_betas; i += blockDim.x) {
            const int f = i / num_betas, k = i % num_betas;
            s_beta[f][k] = f < nf ? beta[(size_t)(f0 + f) * num_betas + k] : 0.f;
        }
        for (int i = threadIdx.x; i < FB * NJ * 12; i += blockDim.x) {
            const int f = i / (NJ * 12);
            s_A[f][i % (NJ * 12)] = f < nf ? A[(size_t)(f0 + f) * NJ * 12 + i % (NJ * 12)] : 0.f;
        }
        __syncthreads();
        float acc[FB];
#pragma unroll
        for (int f = 0; f < FB; f++) acc[f] = 0.f;
        if (live) {
#pragma unroll 4
            for (int k = ks; k < NPF; k += KS) {
                const float pd = __ldg(posedirs + (size_t)k * VC + gc);
#pragma unroll
                for (int f = 0; f < FB; f++) acc[f] = fmaf(s_pf[f][k], pd, acc[f]);
            }
        }
#pragma unroll
        for (int f = 0; f < FB; f++) s_red[ks][f][c] = acc[f];
        __syncthreads();
        if (ks == 0 && live) {
            // v_shaped = beta . shapedirs + v_template ;  v_posed = pose blend + v_shaped   (batch_smpl.py:312, :335)
            float sh[FB];
#pragma unroll
            for (int f = 0; f < FB; f++) sh[f] = 0.f;
            for (int k = 0; k < num_betas; k++) {
                const float sd = __ldg(shapedirs + (size_t)k * VC + gc);
#pragma unroll
                for (int f = 0; f < FB; f++) sh[f] = fmaf(s_beta[f][k], sd, sh[f]);
            }
            const float vt = __ldg(v_template + gc);
#pragma unroll
            for (int f = 0; f < FB; f++) {
                const float pose = (s_red[0][f][c] + s_red[1][f][c]) + (s_red[2][f][c] + s_red[3][f][c]);
                s_red[0][f][c] = pose + (sh[f] + vt);
            }
        }
        __syncthreads();
        // skinning (batch_smpl.py:343-367): T = sum_j w[v][j] A[j],  out = T [v_posed; 1]
        for (int i = threadIdx.x; i < VB * nf; i += blockDim.x) {
            const int lv = i % VB, f = i / VB;
            const int v = v0 + lv;
            if (v >= V) continue;
            float T[12];
#pragma unroll
            for (int e = 0; e < 12; e++) T[e] = 0.f;
            const float* wr = weights + (size_t)v * NJ;
#pragma unroll 4
            for (int j = 0; j < NJ; j++) {
                const float wj = __ldg(wr + j);
#pragma unroll
                for (int e = 0; e < 12; e++) T[e] = fmaf(wj, s_A[f][j * 12 + e], T[e]);
            }
            const float x = s_red[0][f][lv * 3], y = s_red[0][f][lv * 3 + 1], z = s_red[0][f][lv * 3 + 2];
            float* o = verts + ((size_t)(f0 + f) * V + v) * 3;
            o[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
            o[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
            o[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
        }
    }
}

__global__ void __launch_bounds__(128) k_smpl_joints(const float* __restrict__ verts, const float* __restrict__ reg_t,
                                                    int V, int nj, float* __restrict__ joints,
                                                    const float* __restrict__ cam, float* __restrict__ j2d)
{
    const int j = blockIdx.x, b = blockIdx.y;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    const float* vb = verts + (size_t)b * V * 3;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        const float w = __ldg(reg_t + (size_t)j * V + v);
        if (w != 0.f) {
            a0 = fmaf(w, vb[v * 3], a0); a1 = fmaf(w, vb[v * 3 + 1], a1); a2 = fmaf(w, vb[v * 3 + 2], a2);
        }
    }
    __shared__ float red[3][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a0; red[1][threadIdx.x >> 5] = a1; red[2][threadIdx.x >> 5] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        const float s = (red[threadIdx.x][0] + red[threadIdx.x][1]) + (red[threadIdx.x][2] + red[threadIdx.x][3]);
        joints[((size_t)b * nj + j) * 3 + threadIdx.x] = s;
        // batch_orth_proj_idrot (batch_smpl.py:221-233): j2d = s * (xy + t)
        if (j2d && threadIdx.x < 2)
            j2d[((size_t)b * nj + j) * 2 + threadIdx.x] = cam[b * 3] * (s + cam[b * 3 + 1 + threadIdx.x]);
    }
}

}  // namespace

extern "C" size_t lwb_smpl_workspace_bytes(int batch)
{
    if (batch <= 0) return 0;
    return (size_t)batch * (NPF + NJ * 12) * sizeof(float);
}

extern "C" int lwb_smpl_forward(const float* beta, const float* theta, int batch, int num_betas, int num_verts,
                                const float* v_template, const float* shapedirs, const float* posedirs,
                                const float* j_template, const float* j_shapedirs, const int* parents,
                                const float* weights, const float* joint_regressor_t, int num_joints, int rotate_base,
                                float* verts, float* joints, float* Rs, float* J_transformed,
                                const float* cam, float* j2d, void* workspace, lwb_stream_t stream)
{
    LWB_CHECK_ARG(beta && theta && v_template && shapedirs && posedirs && j_template && j_shapedirs && parents && weights,
                  "null model / input pointer");
    LWB_CHECK_ARG(verts && workspace, "null output / workspace pointer");
    LWB_CHECK_ARG(batch > 0 && num_verts > 0, "batch and num_verts must be positive");
    LWB_CHECK_ARG(num_betas > 0 && num_betas <= MAX_BETAS, "num_betas must be in [1,16]");
    LWB_CHECK_ARG(!joints || (joint_regressor_t && num_joints > 0), "joints requested without a regressor");
    LWB_CHECK_ARG(!j2d || (joints && cam), "j2d needs joints and cam");
    cudaStream_t st = (cudaStream_t)stream;
    float* pf = (float*)workspace;
    float* A = pf + (size_t)batch * NPF;
    k_smpl_chain<<<batch, 32, 0, st>>>(beta, theta, num_betas, j_template, j_shapedirs, parents, rotate_base,
                                       pf, A, Rs, J_transformed);
    LWB_LAUNCH_OK();
    k_smpl_skin<<<lwb::ceil_div(num_verts, VB), VB * 3 * KS, 0, st>>>(beta, num_betas, batch, num_verts, v_template,
                                                                       shapedirs, posedirs, weights, pf, A, verts);
    LWB_LAUNCH_OK();
    if (joints) {
        k_smpl_joints<<<dim3(num_joints, batch), 128, 0, st>>>(verts, joint_regressor_t, num_verts, num_joints, joints, cam, j2d);
        LWB_LAUNCH_OK();
    }
    return LWB_OK;
}
