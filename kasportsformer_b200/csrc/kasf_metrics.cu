// kasf_metrics.cu -- K8: GPU evaluation epilogue (one thread per frame, fp64) + joint flip.
//
// Replaces the per-clip numpy loop of reference train_and_evaluate_sp.py:55-103 and the metric
// functions of utils/error_calc.py:5-48:
//   flip-TTA average (fp32, like the reference does on device, :46-51) -> joint 0 := 0 (:55)
//   -> de-normalise xy/z by res_w,res_h (:63-66) -> x factor[t] (:68-70) -> root-relative (:71-72)
//   -> MPJPE, per-joint error, acceleration error, Procrustes-MPJPE (3x3 SVD by one-sided Jacobi)
//   -> per-action partial sums (the means of :105-127 are taken by the host after the cross-rank
//      reduction).
// HBM-bound and tiny (2 x 204 B per frame); shared-memory privatised fp64 accumulators.
#include "kasf_internal.h"

namespace kasf {

__constant__ int c_flip[17] = KASF_FLIP;

__global__ void flip_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / 51;
        const int e = (int)(i % 51), j = e / 3, c = e % 3;
        const float v = in[f * 51 + c_flip[j] * 3 + c];
        out[i] = c == 0 ? -v : v;
    }
}

int launch_flip(const float* in, float* out, long long frames, cudaStream_t st) {
    if (frames <= 0) return KASF_OK;
    const long long n = frames * 51;
    flip_kernel<<<(int)min((n + 255) / 256, (long long)sm_count() * 16), 256, 0, st>>>(in, out, n);
    return cuda_status();
}

// denormalised, root-relative prediction of one frame
__device__ void load_pred(const float* pred, const float* pflip, long long frame, double w, double h, double factor,
                          double (&p)[17][3]) {
    const float* a = pred + frame * 51;
    for (int j = 0; j < 17; ++j) {
        float v[3] = {a[j * 3], a[j * 3 + 1], a[j * 3 + 2]};
        if (pflip) {
            const float* q = pflip + frame * 51 + c_flip[j] * 3;
            v[0] = (v[0] + (-q[0])) / 2.0f;
            v[1] = (v[1] + q[1]) / 2.0f;
            v[2] = (v[2] + q[2]) / 2.0f;
        }
        if (j == 0) v[0] = v[1] = v[2] = 0.f;
        p[j][0] = ((double)v[0] + 1.0) * w / 2 * factor;
        p[j][1] = ((double)v[1] + h / w) * w / 2 * factor;
        p[j][2] = (double)v[2] * w / 2 * factor;
    }
    for (int j = 16; j >= 0; --j)
        for (int c = 0; c < 3; ++c) p[j][c] -= p[0][c];
}
__device__ void load_gt(const float* gt, long long frame, double (&g)[17][3]) {
    const float* a = gt + frame * 51;
    for (int j = 0; j < 17; ++j)
        for (int c = 0; c < 3; ++c) g[j][c] = (double)a[j * 3 + c];
    for (int j = 16; j >= 0; --j)
        for (int c = 0; c < 3; ++c) g[j][c] -= g[0][c];
}

// H = U diag(s) V^T by one-sided Jacobi; singular values sorted descending. A holds U*diag(s) on exit.
__host__ __device__ void svd3(double (&A)[3][3], double (&V)[3][3], double (&s)[3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double al = 0, be = 0, ga = 0;
                for (int i = 0; i < 3; ++i) {
                    al += A[i][p] * A[i][p];
                    be += A[i][q] * A[i][q];
                    ga += A[i][p] * A[i][q];
                }
                if (fabs(ga) <= 1e-300 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
                off = fmax(off, fabs(ga) / sqrt(al * be));
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int i = 0; i < 3; ++i) {
                    const double ap = A[i][p], aq = A[i][q];
                    A[i][p] = c * ap - sn * aq;
                    A[i][q] = sn * ap + c * aq;
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - sn * vq;
                    V[i][q] = sn * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    for (int k = 0; k < 3; ++k) s[k] = sqrt(A[0][k] * A[0][k] + A[1][k] * A[1][k] + A[2][k] * A[2][k]);
    // sort descending (3 elements)
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2 - a; ++b)
            if (s[b] < s[b + 1]) {
                double t = s[b]; s[b] = s[b + 1]; s[b + 1] = t;
                for (int i = 0; i < 3; ++i) {
                    t = A[i][b]; A[i][b] = A[i][b + 1]; A[i][b + 1] = t;
                    t = V[i][b]; V[i][b] = V[i][b + 1]; V[i][b + 1] = t;
                }
            }
}

__host__ __device__ double det3(const double (&M)[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

// Procrustes-aligned MPJPE of one frame (reference utils/error_calc.py:21-48; X = target, Y = predict)
__host__ __device__ double p_mpjpe_frame(const double (&p)[17][3], const double (&g)[17][3]) {
    double muX[3] = {0, 0, 0}, muY[3] = {0, 0, 0};
    for (int j = 0; j < 17; ++j)
        for (int c = 0; c < 3; ++c) muX[c] += g[j][c], muY[c] += p[j][c];
    for (int c = 0; c < 3; ++c) muX[c] /= 17.0, muY[c] /= 17.0;
    double nX = 0, nY = 0;
    for (int j = 0; j < 17; ++j)
        for (int c = 0; c < 3; ++c) {
            const double a = g[j][c] - muX[c], b = p[j][c] - muY[c];
            nX += a * a, nY += b * b;
        }
    nX = sqrt(nX), nY = sqrt(nY);
    double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // H = X0^T Y0
    for (int j = 0; j < 17; ++j)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) H[a][b] += ((g[j][a] - muX[a]) / nX) * ((p[j][b] - muY[b]) / nY);
    double V[3][3], s[3];
    svd3(H, V, s);                                          // H now holds U*diag(s)
    double U[3][3];
    for (int k = 0; k < 3; ++k) {
        const double inv = s[k] > 1e-300 ? 1.0 / s[k] : 0.0;
        for (int i = 0; i < 3; ++i) U[i][k] = H[i][k] * inv;
    }
    if (s[2] <= 1e-14 * s[0]) {   // rank deficient: complete U with a cross product (sign fixed by det below)
        U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
        U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
        U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
    // R = V U^T ; reflection fix: flip last column of V and last singular value
    const double d = det3(V) * det3(U);
    const double sg = d > 0 ? 1.0 : (d < 0 ? -1.0 : 0.0);
    for (int i = 0; i < 3; ++i) V[i][2] *= sg;
    const double tr = s[0] + s[1] + s[2] * sg;
    double R[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = V[i][0] * U[j][0] + V[i][1] * U[j][1] + V[i][2] * U[j][2];
    const double a = tr * nX / nY;
    double t[3];
    for (int c = 0; c < 3; ++c) t[c] = muX[c] - a * (muY[0] * R[0][c] + muY[1] * R[1][c] + muY[2] * R[2][c]);
    double e = 0;
    for (int j = 0; j < 17; ++j) {
        double q = 0;
        for (int c = 0; c < 3; ++c) {
            const double al = a * (p[j][0] * R[0][c] + p[j][1] * R[1][c] + p[j][2] * R[2][c]) + t[c] - g[j][c];
            q += al * al;
        }
        e += sqrt(q);
    }
    return e / 17.0;
}

constexpr int MAX_SMEM_ACTIONS = 64;

__global__ void __launch_bounds__(128)
metrics_kernel(int T, const float* __restrict__ pred, const float* __restrict__ pflip, const float* __restrict__ gt,
               const float* __restrict__ res, const float* __restrict__ factor, const int32_t* __restrict__ action,
               int n_actions, double* __restrict__ sums, double* __restrict__ per_frame, long long frames) {
    __shared__ double s_acc[MAX_SMEM_ACTIONS * KASF_METRIC_COLS];
    const bool priv = n_actions <= MAX_SMEM_ACTIONS;
    if (priv) {
        for (int i = threadIdx.x; i < n_actions * KASF_METRIC_COLS; i += blockDim.x) s_acc[i] = 0.0;
        __syncthreads();
    }
    double* acc = priv ? s_acc : sums;
    for (long long fr = (long long)blockIdx.x * blockDim.x + threadIdx.x; fr < frames;
         fr += (long long)gridDim.x * blockDim.x) {
        const long long b = fr / T;
        const int t = (int)(fr % T);
        const double w = res[b * 2], h = res[b * 2 + 1];
        double p[17][3], g[17][3];
        const int act = action ? action[b] : 0;
        // an action index outside [0, n_actions) is a caller error: the clip is NOT accumulated (it must not corrupt
        // another action's means); column 3 then sums to fewer than B * T frames, which finalize_metrics checks
        if (act < 0 || act >= n_actions) {
            if (per_frame) per_frame[fr * 3] = per_frame[fr * 3 + 1] = per_frame[fr * 3 + 2] = nan("");
            continue;
        }
        load_pred(pred, pflip, fr, w, h, factor[fr], p);
        load_gt(gt, fr, g);
        double* row = acc + act * KASF_METRIC_COLS;
        double e1 = 0;
        for (int j = 0; j < 17; ++j) {
            const double dx = p[j][0] - g[j][0], dy = p[j][1] - g[j][1], dz = p[j][2] - g[j][2];
            const double e = sqrt(dx * dx + dy * dy + dz * dz);
            e1 += e;
            atomicAdd(row + 5 + j, e);
        }
        e1 /= 17.0;
        const double e2 = p_mpjpe_frame(p, g);
        atomicAdd(row + 0, e1);
        atomicAdd(row + 1, e2);
        atomicAdd(row + 3, 1.0);
        double ea = 0.0;
        if (t + 2 < T) {    // acceleration error (utils/error_calc.py:15-19)
            double p1[17][3], g1[17][3], p2[17][3], g2[17][3];
            load_pred(pred, pflip, fr + 1, w, h, factor[fr + 1], p1);
            load_gt(gt, fr + 1, g1);
            load_pred(pred, pflip, fr + 2, w, h, factor[fr + 2], p2);
            load_gt(gt, fr + 2, g2);
            for (int j = 0; j < 17; ++j) {
                double q = 0;
                for (int c = 0; c < 3; ++c) {
                    const double d = (p[j][c] - 2 * p1[j][c] + p2[j][c]) - (g[j][c] - 2 * g1[j][c] + g2[j][c]);
                    q += d * d;
                }
                ea += sqrt(q);
            }
            ea /= 17.0;
            atomicAdd(row + 2, ea);
            atomicAdd(row + 4, 1.0);
        }
        if (per_frame) {
            per_frame[fr * 3] = e1;
            per_frame[fr * 3 + 1] = e2;
            per_frame[fr * 3 + 2] = ea;
        }
    }
    if (priv) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_actions * KASF_METRIC_COLS; i += blockDim.x)
            if (s_acc[i] != 0.0) atomicAdd(sums + i, s_acc[i]);
    }
}

// host-side execution of the very same Procrustes routine (self-test hook, CPU unit tests)
double host_p_mpjpe(const double* p, const double* g) {
    double P[17][3], G[17][3];
    for (int j = 0; j < 17; ++j)
        for (int c = 0; c < 3; ++c) P[j][c] = p[j * 3 + c], G[j][c] = g[j * 3 + c];
    return p_mpjpe_frame(P, G);
}

int launch_metrics(int T, const float* pred, const float* pred_flip, const float* gt, const float* res,
                   const float* factor, const int32_t* action, int n_actions, double* sums, double* per_frame,
                   int B, cudaStream_t st) {
    if (B <= 0) return KASF_OK;
    if (T < 1 || n_actions < 1) return KASF_EINVAL;
    const long long frames = (long long)B * T;
    const int grid = (int)min((frames + 127) / 128, (long long)sm_count() * 8);
    metrics_kernel<<<grid, 128, 0, st>>>(T, pred, pred_flip, gt, res, factor, action, n_actions, sums, per_frame, frames);
    return cuda_status();
}

}  // namespace kasf
