/* TEST INFRASTRUCTURE -- an oracle-backed stand-in for the subset of include/fgoicp_c.h that the C++ host driver
 * (fast_go_icp_b200/csrc/fgoicp_host.cpp) calls, so that the drop-in class icp::FastGoICP -- its preprocessing, the
 * level-synchronous and best-first outer searches, wave scheduling, pruning, result restoration -- can be exercised on a
 * machine without a GPU.  Every entry point answers with the CPU oracle (oracle/libfgoicp_oracle.so) what the CUDA
 * library answers with kernels; semantics follow csrc/bnb.cu and csrc/nn_icp.cu entry point by entry point.
 * Linked ONLY into build/fgoicp_harness_cpu by tests/test_cpp_host_cpu.py; never part of libfgoicp_b200.so. */
#include <fgoicp_c.h>

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* the oracle keeps process-wide state (k-d tree cache, trim switch): calls from the host driver's per-device threads
 * are serialised here -- the stand-in checks the driver's sharding logic, not concurrency of the oracle */
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_contexts_created = 0;

/* oracle exports (oracle/fgoicp_oracle.c) */
void orc_lut_dims(const float* bbox_min, const float* bbox_max, float res, int* dims);
void orc_lut_build(const float* model, size_t nt, const float* bbox_min, float res, const int* dims, float* out);
float orc_rotation(float x, float y, float z, float* R);
float orc_icp(const float* model, size_t nt, const float* data, size_t ns, int max_iter, float thr, const float* R0,
              const float* t0, float* Rout, float* tout, int* iters_out);
float orc_bnb_r3(const float* model, size_t nt, const float* data, size_t ns, const float* lut, const int* dims,
                 const float* bbox_min, float res, const float* rot_xyz_span, int fix_rot, float best_sse,
                 float sse_threshold, int batch, float* best_t, uint64_t* evals, uint32_t* batches);
void orc_set_trim_k(size_t k);
size_t orc_trim_count(size_t ns, float rho);

struct fgoicp_ctx
{
    float *model, *data, *lut;
    size_t nt, ns;
    int dims[3];
    float bbox_min[3], res;
};

static char g_err[256] = "";
const char* fgoicp_last_error(void) { return g_err; }
const char* fgoicp_version(void) { return "oracle-backed test stand-in (no CUDA)"; }

int fgoicp_ctx_create(const float* model_xyz, size_t nt, const float* data_xyz, size_t ns, const float bbox_min[3],
                      const float bbox_max[3], float lut_resolution, int device, unsigned flags, fgoicp_ctx** out)
{
    fgoicp_ctx* c;
    size_t cells;
    (void)flags;
    if (!out || !model_xyz || !data_xyz || nt == 0 || ns == 0 || !(lut_resolution > 0.0f))
    {
        snprintf(g_err, sizeof(g_err), "bad argument");
        return FGOICP_ERR_ARG;
    }
    pthread_mutex_lock(&g_lock);
    ++g_contexts_created;
    if (getenv("ORACLE_ABI_TRACE")) fprintf(stderr, "oracle_abi: context %d on device %d\n", g_contexts_created, device);
    c = (fgoicp_ctx*)calloc(1, sizeof(*c));
    c->nt = nt; c->ns = ns; c->res = lut_resolution;
    c->model = (float*)malloc(sizeof(float) * 3 * nt); memcpy(c->model, model_xyz, sizeof(float) * 3 * nt);
    c->data = (float*)malloc(sizeof(float) * 3 * ns); memcpy(c->data, data_xyz, sizeof(float) * 3 * ns);
    memcpy(c->bbox_min, bbox_min, sizeof(c->bbox_min));
    orc_lut_dims(bbox_min, bbox_max, lut_resolution, c->dims);
    cells = (size_t)c->dims[0] * c->dims[1] * c->dims[2];
    c->lut = (float*)malloc(sizeof(float) * cells);
    orc_lut_build(c->model, nt, bbox_min, lut_resolution, c->dims, c->lut);
    *out = c;
    pthread_mutex_unlock(&g_lock);
    return FGOICP_OK;
}

int fgoicp_ctx_destroy(fgoicp_ctx* c)
{
    if (!c) return FGOICP_OK;
    orc_set_trim_k(0);
    free(c->model); free(c->data); free(c->lut); free(c);
    return FGOICP_OK;
}

int fgoicp_ctx_info(const fgoicp_ctx* c, fgoicp_info* o)
{
    memset(o, 0, sizeof(*o));
    o->nt = c->nt; o->ns = c->ns; o->resolution = c->res;
    o->dims[0] = c->dims[0]; o->dims[1] = c->dims[1]; o->dims[2] = c->dims[2];
    return FGOICP_OK;
}

int fgoicp_set_sampler(fgoicp_ctx* c, int sampler) { (void)c; (void)sampler; return FGOICP_OK; }

int fgoicp_set_trim(fgoicp_ctx* c, float trim_fraction, uint64_t* inliers)
{
    size_t k = orc_trim_count(c->ns, trim_fraction);
    orc_set_trim_k(k == c->ns ? 0 : k);
    if (inliers) *inliers = k;
    return FGOICP_OK;
}

int fgoicp_preprocess(float* model_xyz, size_t nt, float* data_xyz, size_t ns, int device, unsigned flags,
                      fgoicp_normalisation* out)
{
    (void)model_xyz; (void)nt; (void)data_xyz; (void)ns; (void)device; (void)flags; (void)out;
    snprintf(g_err, sizeof(g_err), "no CUDA device available: this library has no CPU fallback");
    return FGOICP_ERR_CUDA;
}

int fgoicp_icp(fgoicp_ctx* c, const float R0[9], const float t0[3], int max_iter, float thr, float* sse, float R[9],
               float t[3], int* iters)
{
    float Ro[9], to[3];
    int it = 0;
    float e;
    pthread_mutex_lock(&g_lock);
    e = orc_icp(c->model, c->nt, c->data, c->ns, max_iter, thr, R0, t0, Ro, to, &it);
    pthread_mutex_unlock(&g_lock);
    if (sse) *sse = e;
    if (R) memcpy(R, Ro, sizeof(Ro));
    if (t) memcpy(t, to, sizeof(to));
    if (iters) *iters = it;
    return FGOICP_OK;
}

static int bnb_r3_unlocked(fgoicp_ctx* c, const float rot_xyz_span[4], int fix_rot, float best_sse, float sse_threshold,
                           float* best_ub, float best_t[3], uint64_t* evals)
{
    uint64_t ev = 0;
    uint32_t nb = 0;
    float bt[3] = { 0, 0, 0 };
    float ub = orc_bnb_r3(c->model, c->nt, c->data, c->ns, c->lut, c->dims, c->bbox_min, c->res, rot_xyz_span, fix_rot,
                          best_sse, sse_threshold, 32, bt, &ev, &nb);
    if (best_ub) *best_ub = ub;
    if (best_t) memcpy(best_t, bt, sizeof(bt));
    if (evals) *evals = ev;
    return FGOICP_OK;
}

int fgoicp_bnb_r3(fgoicp_ctx* c, const float rot_xyz_span[4], int fix_rot, float best_sse, float sse_threshold,
                  float* best_ub, float best_t[3], uint64_t* evals)
{
    int rc;
    pthread_mutex_lock(&g_lock);
    rc = bnb_r3_unlocked(c, rot_xyz_span, fix_rot, best_sse, sse_threshold, best_ub, best_t, evals);
    pthread_mutex_unlock(&g_lock);
    return rc;
}

/* csrc/bnb.cu fgoicp_so3_level_ub: fixed-rotation searches of every cube against the level-start best_sse, then ICP on
 * the cubes with ub < 1.8 * best_sse (double compare), winners taken in ascending cube order */
int fgoicp_so3_level_ub(fgoicp_ctx* c, const float* cubes, int n, float best_sse, float sse_threshold, float* ub, float* bt,
                        float* io_best_sse, float io_best_R[9], float io_best_t[3], fgoicp_level_stats* stats)
{
    int i;
    if (stats) { memset(stats, 0, sizeof(*stats)); stats->best_icp_index = -1; }
    pthread_mutex_lock(&g_lock);
    for (i = 0; i < n; ++i)
    {
        uint64_t ev = 0;
        bnb_r3_unlocked(c, cubes + 4 * i, 1, best_sse, sse_threshold, &ub[i], bt + 3 * i, &ev);
        if (stats) stats->evals += ev;
    }
    for (i = 0; i < n; ++i)
    {
        float R0[9], R[9], t[3], e;
        int it = 0;
        if (!((double)ub[i] < (double)best_sse * 1.8)) continue;
        orc_rotation(cubes[4 * i], cubes[4 * i + 1], cubes[4 * i + 2], R0);
        e = orc_icp(c->model, c->nt, c->data, c->ns, 100, (float)0.005, R0, bt + 3 * i, R, t, &it);
        if (stats) { stats->n_icp += 1; stats->icp_iters += (uint32_t)it; }
        if (e < *io_best_sse)
        {
            *io_best_sse = e;
            memcpy(io_best_R, R, sizeof(R)); memcpy(io_best_t, t, sizeof(t));
            if (stats) stats->best_icp_index = i;
        }
    }
    pthread_mutex_unlock(&g_lock);
    return FGOICP_OK;
}

int fgoicp_so3_level_lb(fgoicp_ctx* c, const float* cubes, int n, float best_sse, float sse_threshold, float* lb,
                        fgoicp_level_stats* stats)
{
    int i;
    if (stats) memset(stats, 0, sizeof(*stats));
    pthread_mutex_lock(&g_lock);
    for (i = 0; i < n; ++i)
    {
        uint64_t ev = 0;
        float dummy[3];
        bnb_r3_unlocked(c, cubes + 4 * i, 0, best_sse, sse_threshold, &lb[i], dummy, &ev);
        if (stats) stats->evals += ev;
    }
    pthread_mutex_unlock(&g_lock);
    return FGOICP_OK;
}
