// Host side of the B200 EqF path: the filter driver of the reference's VIOFilter
// (src/VIOFilter.cpp) re-bodied over device-resident state, exported through the C ABI declared in
// include/eqvio_b200.h.  Everything numerical runs in the kernels of kernels.cuh; the host keeps
// only the IMU buffer, the landmark id list and the discrete bookkeeping decisions
// (which landmarks to drop / add), made in fp64 from per-landmark scalars the gate kernel returns.
//
// There is no CPU fallback: every entry point that touches filter state needs a CUDA device and
// returns EQVIO_ERR_CUDA otherwise.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/eqvio_b200.h"
#include "kernels.cuh"
#include "structured_riccati.cuh"
#include "blockchol.cuh"

using namespace eqvio;

#define EQVIO_STR2(x) #x
#define EQVIO_STR(x) EQVIO_STR2(x)

namespace {

std::string g_createError;

struct ImuSample {
    double stamp;
    double v[12];  // gyr, acc, gyrBiasVel, accBiasVel
};

enum { PROF_PROP_LL = 0, PROF_PANEL, PROF_TRAIL, PROF_SYRK, PROF_BC_DIAG, PROF_BC_PANEL, PROF_BC_TRAIL, PROF_CLASSES };
static_assert(PROF_CLASSES == EQVIO_PROF_CLASSES, "profile classes of the header");

struct EventPair {
    cudaEvent_t a, b;
    int cls;
};

}  // namespace

struct eqvio_filter {
    eqvio_settings st;
    int device = 0, cap = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    std::string err;

    // host-side filter state
    bool initialised = false;
    double time = -1.0;
    std::deque<ImuSample> buf;
    std::vector<int> ids;  // state order
    double xi0s[23];       // host mirror of xi0.sensor
    std::vector<int> lastOutliers;

    // device state
    int ld = 0;
    double* Sig[2] = {nullptr, nullptr};
    CUtensorMap sigMap[2] = {};  // 2-D tensor maps of the two covariance buffers (96 x 96 box): TMA staging of the S blocks
    bool haveSigMap = false;
    int cur = 0;
    double* lm[2] = {nullptr, nullptr};
    int* dids[2] = {nullptr, nullptr};
    int lmcur = 0;
    double* d_xi0s = nullptr;
    double* d_Xs[2] = {nullptr, nullptr};  // X sensor part, ping-pong across the observer integration
    int xcur = 0;
    cudaStream_t stream4 = nullptr, stream5 = nullptr;  // block sweep: next-column trailing tiles / the urgent pair
    cudaStream_t stream3 = nullptr;  // low priority: deferred downdate tiles of the look-ahead correction
    cudaStream_t stream2 = nullptr;  // observer chain of the propagation runs beside the Riccati chain
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    RiccatiCtx* d_ctx = nullptr;
    ObsStep* d_steps = nullptr;
    // per-frame input block (FrameHeader | imu | y | measIdx | lmOf) at fixed device / pinned addresses
    unsigned char* d_frame = nullptr;
    unsigned char* h_frame = nullptr;
    unsigned char* hd_frame = nullptr;  // device-side address of the pinned frame block (zero-copy upload kernel)
    unsigned char* hd_out = nullptr;    // ... of the pinned result block
    int zeroCopy = 1;                   // frame / result blocks move through block_copy_kernel instead of memcpy nodes
    int prefetchSigma = 1;              // the frame upload kernel also prefetches the covariance into L2
    int propFusion = 1;                 // fast Riccati step: prop_ll_kernel builds its own factors, the strip runs beside it
    cudaEvent_t evStrip0 = nullptr, evStrip1 = nullptr;
    int earlyRows = 1;                  // steady block-sweep frames: measurement rows right behind the observer, gate beside the sweep
    int measHookNm = 0;                 // > 0: enqueue_propagation launches the measurement rows on the observer's stream
    cudaEvent_t evG0 = nullptr, evGate = nullptr;
    int splitDowndate = 1;              // two CTAs per Sigma tile when the whole downdate is a single wave
    size_t frameBytes = 0, offImu = 0, offY = 0, offMeasIdx = 0, offLmOf = 0, offYIdx = 0;
    int* d_yIdx = nullptr;
    FrameHeader* d_hdrSteps = nullptr;  // one header per IMU sample (per-sample Riccati variants)
    // compact blocks of the per-sample / Normal-chart Riccati variants (structured_riccati.cuh), allocated at first use
    struct Sric {
        double *Ts = nullptr, *Tl = nullptr, *Es = nullptr, *El = nullptr, *uv = nullptr, *ladder = nullptr, *dtBs = nullptr;
        unsigned long long* norm = nullptr;
        SE3* cc = nullptr;  // camera-frame changes of the 43 sensor evaluations of stateMatrixADiscrete
        void release() {
            cudaFree(Ts); cudaFree(Tl); cudaFree(Es); cudaFree(El); cudaFree(uv); cudaFree(ladder); cudaFree(dtBs); cudaFree(norm); cudaFree(cc);
            Ts = Tl = Es = El = uv = ladder = dtBs = nullptr;
            norm = nullptr;
            cc = nullptr;
        }
    } sric;
    FrameHeader* d_hdr = nullptr;
    double* d_imu = nullptr;
    int maxSteps = 0;
    int yCap = 0;
    size_t offKeepF = 0, offNewMeasF = 0, offMapF = 0, offNewIdsF = 0, offCounts = 0;
    int *d_keepF = nullptr, *d_newMeasF = nullptr, *d_mapF = nullptr, *d_newIdsF = nullptr, *d_counts = nullptr;
    // fixed pinned output block of the steady path: gate scalars | spec flag | status words
    unsigned char* h_out = nullptr;
    size_t outOffSpec = 0, outOffStatus = 0, outOffEst = 0, outBytes = 0;
    unsigned char* d_mapblk = nullptr;  // d_newP | d_map | d_newIds
    size_t mapBlkBytes = 0;
    unsigned char* d_outblk = nullptr;  // device mirror of h_out: d_gate / d_spec / d_status / d_out point into it
    bool estValid = false;  // h_out holds the state estimate of the current state (produced by the steady update)
    // CUDA graphs of the steady-state update, keyed by everything that shapes the launch sequence
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        int cur2 = 0, lmcur2 = 0, xcur2 = 0;
        long long launches = 0;
        unsigned long long lastUse = 0;
    };
    std::map<std::vector<int>, GraphEntry> graphs;
    unsigned long long graphClock = 0;
    int useGraph = 1;
    long long graphLaunches = 0, graphCaptures = 0;
    double updateMs = 0;
    double *d_rows = nullptr, *d_uv = nullptr, *d_Z = nullptr, *d_Lout = nullptr;
    size_t zElems = 0;
    double *d_Gamma2 = nullptr, *d_ytilde = nullptr;
    int corrMode = 2;    // 0: sequential chunks, 1: batch Cholesky sweep over Z, 2 (default): block sweep with look-ahead (blockchol.cuh) up to BC_MAX_ROWS measurement rows, sequential chunks beyond
    int* d_bcCnt = nullptr;  // mode 2: per trailing step, the number of its CTAs that have finished
    double *d_bcZ = nullptr, *d_bcZp = nullptr, *d_bcMt = nullptr;  // mode 2: augmented matrix, panel tiles of the current block column, LT | XT per block (blockchol.cuh)
    std::vector<cudaEvent_t> bcEv;
    int speculate = 1;   // launch the correction before the gate results reach the host (redone on a gate hit)
    int downdateTC = 0;  // 1: tcgen05 split-bf16 downdate (fp32 accumulate in TMEM) instead of the fp64 DMMA one
    unsigned char* d_Ysplit = nullptr;
    int lazyMirror = 1;  // downdate refreshes the upper triangle only where the next chunk reads it
    std::vector<int> h_lmOfSorted;  // state indices of the correction rows (host copy of d_lmOf)
    double* d_normalM = nullptr;  // Normal chart: sensor block of the coordinate differential and its inverse (2 x 441)
    bool normalMValid = false;    // ... which only depends on xi0's sensor part: recomputed when that is set
    int prLeast = 0, prGreatest = 0;  // stream priority range of the device
    int smCount = 148;                // SMs of the device (grid sizing)
    bool pdlHold = false;  // next launch_pdl is a plain launch (its predecessor produces what the kernel reads before its wait)
    int pdl = 1;           // chunk kernels are launched with programmatic dependent launch allowed
    int specNew = 1;       // frames with new ids also speculate (new-landmark positions computed on the device)
    int stageS = 1;        // chunk factor kernel: Sigma[L_c, L_c] as one TMA tensor copy when the chunk is contiguous in the state
    int *d_keepI = nullptr, *d_newMeas = nullptr;
    int newMeasCap = 0;
    int fuseSmall = 1;     // steady update: gate + measurement rows in one launch, lift + state estimate in one launch
    int fuseObserver = 1;  // sensor + landmark parts of the observer integration as one software-pipelined kernel
    const char* tlNames[512] = {nullptr};
    int tlNext = 0;  // debug timeline slot counter (EQVIO_TIMELINE builds)
    double hostUs[4] = {0, 0, 0, 0};  // process_vision host time: phase A (plan + enqueue), phase B, wait for the device, rest of phase C
    long long hostCalls = 0;
    double hostWaitUs = 0;
    int lookahead = 2;   // 2 = automatic (on when the tile grid spans more than one wave, T >= 24), 1 = on, 0 = off: split each downdate into the tiles the next chunk gathers (urgent) and the rest (beside the next factor)
    double* d_Y2 = nullptr;
    // lazy trailing updates of the look-ahead correction (EQVIO_TUNE_LAZY_DOWNDATE = M >= 1): the panels of every chunk stay in d_Z, the
    // deferred tiles are visited once per M chunks (K = 64 M rows of Y per visit), d_lvl says through which chunk a tile is current
    int lazyMerge = 1;
    int bandSplit = 1;   // urgent tiles of the lazy form: two CTAs per tile
    int restPersist = 0;    // 2 = deferred launch of the lazy form as a persistent kernel (two CTAs per SM drawing tiles from a counter)
    int restAfterBand = 1;  // the deferred launch waits for the urgent one: the next factor (a programmatic dependent of the urgent
                            // launch) is resident before the deferred tiles fill the SMs
    int* d_lvl = nullptr;
    std::vector<cudaEvent_t> chunkEv;
    int* d_spec = nullptr;  // [0] set by the gate kernel when any measured landmark exceeds a threshold, [1] constant 0
    cudaEvent_t augEv[2] = {nullptr, nullptr};
    bool augTimed = false;
    int chunkLm = 32;    // landmarks per chunk (<= CH_R / 2)
    double *d_Cblk = nullptr, *d_Gamma = nullptr, *d_gate = nullptr, *d_y = nullptr, *d_newP = nullptr, *d_out = nullptr;
    int *d_measIdx = nullptr, *d_lmOf = nullptr, *d_map = nullptr, *d_newIds = nullptr, *d_status = nullptr;

    // pinned staging: NA arena sets used in rotation, one per API call, so that a call may return while its
    // async copies are still in flight (the set is waited for, through its event, before it is reused)
    static constexpr int NA = 4;
    struct ArenaSet {
        std::vector<std::pair<char*, size_t>> blocks;
        size_t used = 0;
        cudaEvent_t ev = nullptr;
        bool pending = false;
    } arenaSets[NA];
    int arenaCur = 0;

    // measurement hooks
    long long launches = 0;
    bool stageTiming = false;
    cudaEvent_t stageEv[4] = {nullptr, nullptr, nullptr, nullptr};
    double stageMs[3] = {0, 0, 0};
    bool capturing = false;    // inside cudaStreamBeginCapture / EndCapture of a steady update
    bool steadySplit = false;  // the last steady update was enqueued with plain launches: its stage events 1 and 2 are recorded
    double augMs = 0;  // device time of augment_landmark_states calls since the last process_vision
    bool profiling = false;
    std::vector<EventPair> evPool;
    size_t evUsed = 0;
    double profMs[PROF_CLASSES] = {};
    long long profLaunches[PROF_CLASSES] = {};

    // scratch of a process_vision call split in phases (batch API)
    struct Pending {
        bool active = false, gated = false;
        int n = 0;
        std::vector<int> mids;
        std::vector<double> my;
        std::vector<int> measIdx;  // per state landmark
        std::vector<char> keep;
        double* h_gate = nullptr;
        int* h_status = nullptr;
        int nStatus = 0;
        Camera cam;
        bool corrected = false;
        bool speculated = false;       // correction launched before the gate results were read
        int specNewCount = 0;          // new landmarks appended speculatively (positions from the device-side median depth)
        bool steady = false;           // phase A enqueued the whole update (possibly as a CUDA graph)
        bool ignoreGate = false;       // featureRetention leaves no room for removals: the gate flag is moot
        std::vector<int> oldIds;       // state ids when the gate ran (gate results are indexed like this)
        std::vector<char> measKept;    // per measurement: survives gating
        int* h_spec = nullptr;
    } pend;
};

namespace {

#define CUDA_TRY(f, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            (f)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                        \
            return EQVIO_ERR_CUDA;                                                                \
        }                                                                                         \
    } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int dimp_of(int N) { return SOFF + 3 * N; }

void* stage_alloc(eqvio_filter* f, size_t bytes) {
    auto& A = f->arenaSets[f->arenaCur];
    bytes = (bytes + 63) & ~size_t(63);
    if (!A.blocks.empty()) {
        auto& a = A.blocks.back();
        if (A.used + bytes <= a.second) {
            void* p = a.first + A.used;
            A.used += bytes;
            return p;
        }
    }
    size_t sz = std::max(bytes, size_t(1) << 20);
    char* p = nullptr;
    if (cudaMallocHost(&p, sz) != cudaSuccess) return nullptr;
    A.blocks.emplace_back(p, sz);
    A.used = bytes;
    return p;
}
// start of an API call: mark the set used by the previous call as in flight, move to the next one
void stage_reset(eqvio_filter* f) {
    {
        auto& prev = f->arenaSets[f->arenaCur];
        if (!prev.ev) cudaEventCreateWithFlags(&prev.ev, cudaEventDisableTiming);
        cudaEventRecord(prev.ev, f->stream);
        prev.pending = true;
    }
    f->arenaCur = (f->arenaCur + 1) % eqvio_filter::NA;
    auto& A = f->arenaSets[f->arenaCur];
    if (A.pending) {
        cudaEventSynchronize(A.ev);
        A.pending = false;
    }
    // keep only the largest block
    while (A.blocks.size() > 1) {
        size_t smallest = 0;
        for (size_t i = 1; i < A.blocks.size(); ++i)
            if (A.blocks[i].second < A.blocks[smallest].second) smallest = i;
        cudaFreeHost(A.blocks[smallest].first);
        A.blocks.erase(A.blocks.begin() + smallest);
    }
    A.used = 0;
}

template <class T>
int upload(eqvio_filter* f, T* dst, const T* src, size_t count) {
    if (count == 0) return EQVIO_OK;
    T* h = static_cast<T*>(stage_alloc(f, count * sizeof(T)));
    if (!h) {
        f->err = "pinned staging allocation failed";
        return EQVIO_ERR_CUDA;
    }
    std::memcpy(h, src, count * sizeof(T));
    CUDA_TRY(f, cudaMemcpyAsync(dst, h, count * sizeof(T), cudaMemcpyHostToDevice, f->stream));
    return EQVIO_OK;
}
template <class T>
int download_async(eqvio_filter* f, T** hostOut, const T* src, size_t count) {
    T* h = static_cast<T*>(stage_alloc(f, std::max<size_t>(count, 1) * sizeof(T)));
    if (!h) {
        f->err = "pinned staging allocation failed";
        return EQVIO_ERR_CUDA;
    }
    if (count) CUDA_TRY(f, cudaMemcpyAsync(h, src, count * sizeof(T), cudaMemcpyDeviceToHost, f->stream));
    *hostOut = h;
    return EQVIO_OK;
}

int check_launch(eqvio_filter* f, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        f->err = std::string(what) + ": " + cudaGetErrorString(e);
        return EQVIO_ERR_CUDA;
    }
    ++f->launches;
#ifdef EQVIO_TIMELINE
    if (f->tlNext > 0 && f->tlNext <= 512) f->tlNames[f->tlNext - 1] = what;  // the slot TL_SLOT handed to this launch
#endif
    return EQVIO_OK;
}
#ifdef EQVIO_TIMELINE
#define TL_SLOT(f) ((f)->tlNext++)
#else
#define TL_SLOT(f) (-1)
#endif
#define LAUNCH_CHECK(f, what)                          \
    do {                                               \
        int rc_ = check_launch((f), (what));           \
        if (rc_ != EQVIO_OK) return rc_;               \
    } while (0)

// Launch with programmatic dependent launch allowed: the grid may be scheduled while its predecessor on the stream is
// still draining; the kernel itself blocks in griddepcontrol.wait (pdl_wait() in kernels.cuh) before it touches global
// memory, so only launch latency / CTA set-up overlaps.  Under stream capture the edge becomes a programmatic graph edge.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(eqvio_filter* f, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (f->pdl && !f->pdlHold) ? 1 : 0;
    // explicit launch priority (that of the stream): a captured kernel node keeps it, so that the block scheduler still
    // prefers the critical-path kernels over the deferred downdate tiles when the update is replayed as a graph
    attr[1].id = cudaLaunchAttributePriority;
    attr[1].val.priority = st == f->stream3 ? f->prLeast : f->prGreatest;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// profiling brackets -------------------------------------------------------------------------
int prof_begin(eqvio_filter* f, int cls) {
    if (!f->profiling) return -1;
    if (f->evUsed == f->evPool.size()) {
        EventPair p;
        cudaEventCreate(&p.a);
        cudaEventCreate(&p.b);
        f->evPool.push_back(p);
    }
    int k = (int)f->evUsed++;
    f->evPool[k].cls = cls;
    cudaEventRecord(f->evPool[k].a, f->stream);
    return k;
}
void prof_end(eqvio_filter* f, int k) {
    if (k >= 0) cudaEventRecord(f->evPool[k].b, f->stream);
}
void prof_collect(eqvio_filter* f) {  // stream must be synchronised
    for (size_t k = 0; k < f->evUsed; ++k) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, f->evPool[k].a, f->evPool[k].b) == cudaSuccess) {
            f->profMs[f->evPool[k].cls] += ms;
            f->profLaunches[f->evPool[k].cls] += 1;
        }
    }
    f->evUsed = 0;
}
void stage_mark(eqvio_filter* f, int i) {
    if (f->stageTiming) cudaEventRecord(f->stageEv[i], f->stream);
}

Camera to_camera(const eqvio_camera* c) {
    Camera k;
    k.model = c->model;
    k.width = c->width;
    k.height = c->height;
    k.ndist = c->ndist;
    k.fx = c->fx;
    k.fy = c->fy;
    k.cx = c->cx;
    k.cy = c->cy;
    for (int i = 0; i < 5; ++i) {
        k.dist[i] = c->dist[i];
        k.inv_dist[i] = c->inv_dist[i];
    }
    return k;
}

// Initial covariance diagonal in the internal (padded) order (VIOFilterSettings.h:208-229).
void initial_diag(const eqvio_settings& s, int N, bool depthVariance, std::vector<double>& d) {
    d.assign(dimp_of(N), 0.0);
    const double v[7] = {s.initialBiasOmegaVariance,      s.initialBiasAccelVariance,     s.initialAttitudeVariance,
                         s.initialPositionVariance,       s.initialVelocityVariance,      s.initialCameraAttitudeVariance,
                         s.initialCameraPositionVariance};
    for (int i = 0; i < SENSOR_DIM; ++i) d[i] = v[i / 3];
    for (int i = 0; i < N; ++i)
        for (int a = 0; a < 3; ++a)
            d[SOFF + 3 * i + a] =
                (a == 2 && depthVariance && s.initialPointDepthVariance > 0) ? s.initialPointDepthVariance : s.initialPointVariance;
}

// (re)allocate the per-frame input block for `steps` IMU segments and `ycap` measured landmarks
int alloc_frame(eqvio_filter* f, int steps, int ycap) {
    for (auto& g : f->graphs)
        if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    f->graphs.clear();
    cudaFree(f->d_frame);
    cudaFree(f->d_steps);
    cudaFree(f->d_hdrSteps);
    f->d_hdrSteps = nullptr;
    if (f->h_frame) cudaFreeHost(f->h_frame);
    f->d_frame = nullptr;
    f->h_frame = nullptr;
    f->d_steps = nullptr;
    f->maxSteps = steps;
    f->yCap = ycap;
    const size_t cap1 = std::max(f->cap, 1);
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    f->offImu = up(sizeof(FrameHeader));
    f->offY = up(f->offImu + (size_t)steps * 13 * sizeof(double));
    f->offMeasIdx = up(f->offY + (size_t)ycap * 2 * sizeof(double));
    f->offLmOf = up(f->offMeasIdx + cap1 * sizeof(int));
    f->offYIdx = up(f->offLmOf + cap1 * sizeof(int));
    // landmark-set change of a planned frame (lost ids pruned, new ids appended inside the update): keep flags of the old state,
    // measurement indices of the new ids, old-index map and ids of the new state, counts
    f->offKeepF = up(f->offYIdx + cap1 * sizeof(int));
    f->offNewMeasF = up(f->offKeepF + cap1 * sizeof(int));
    f->offMapF = up(f->offNewMeasF + (size_t)std::max(ycap, 1) * sizeof(int));
    f->offNewIdsF = up(f->offMapF + cap1 * sizeof(int));
    f->offCounts = up(f->offNewIdsF + cap1 * sizeof(int));
    f->frameBytes = up(f->offCounts + 64);
    CUDA_TRY(f, cudaMalloc(&f->d_frame, f->frameBytes));
    CUDA_TRY(f, cudaMallocHost(&f->h_frame, f->frameBytes));
    std::memset(f->h_frame, 0, f->frameBytes);
    if (cudaHostGetDevicePointer((void**)&f->hd_frame, f->h_frame, 0) != cudaSuccess) {
        cudaGetLastError();
        f->hd_frame = nullptr;
    }
    CUDA_TRY(f, cudaMalloc(&f->d_steps, (size_t)steps * sizeof(ObsStep)));
    CUDA_TRY(f, cudaMalloc(&f->d_hdrSteps, (size_t)steps * sizeof(FrameHeader)));
    f->d_hdr = reinterpret_cast<FrameHeader*>(f->d_frame);
    f->d_imu = reinterpret_cast<double*>(f->d_frame + f->offImu);
    f->d_y = reinterpret_cast<double*>(f->d_frame + f->offY);
    f->d_measIdx = reinterpret_cast<int*>(f->d_frame + f->offMeasIdx);
    f->d_lmOf = reinterpret_cast<int*>(f->d_frame + f->offLmOf);
    f->d_yIdx = reinterpret_cast<int*>(f->d_frame + f->offYIdx);
    f->d_keepF = reinterpret_cast<int*>(f->d_frame + f->offKeepF);
    f->d_newMeasF = reinterpret_cast<int*>(f->d_frame + f->offNewMeasF);
    f->d_mapF = reinterpret_cast<int*>(f->d_frame + f->offMapF);
    f->d_newIdsF = reinterpret_cast<int*>(f->d_frame + f->offNewIdsF);
    f->d_counts = reinterpret_cast<int*>(f->d_frame + f->offCounts);
    return EQVIO_OK;
}

// CUtensorMap of a column-major ld x ld fp64 covariance buffer with a 96 x 96 box (the S block of one chunk).  The encoder lives in
// the driver library; it is fetched through the runtime so that nothing links against libcuda directly.
bool make_sigma_maps(eqvio_filter* f) {
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)f->ld, (cuuint64_t)f->ld};
    const cuuint64_t strides[1] = {(cuuint64_t)f->ld * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)CH_STG_N, (cuuint32_t)CH_STG_N}, estr[2] = {1, 1};
    for (int k = 0; k < 2; ++k)
        if (reinterpret_cast<EncodeTiled>(fn)(&f->sigMap[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, f->Sig[k], dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    return true;
}

int alloc_sric(eqvio_filter* f);
int alloc_device(eqvio_filter* f) {
    const int cap = f->cap;
    const int dimpMax = dimp_of(cap);
    f->ld = (dimpMax + 127) & ~127;  // whole 64x64 (DMMA) / 128x128 (tcgen05) tiles are moved by the downdate kernels
    const size_t sigElems = (size_t)f->ld * f->ld;
    for (int k = 0; k < 2; ++k) {
        CUDA_TRY(f, cudaMalloc(&f->Sig[k], sigElems * sizeof(double)));
        CUDA_TRY(f, cudaMemsetAsync(f->Sig[k], 0, sigElems * sizeof(double), f->stream));
        CUDA_TRY(f, cudaMalloc(&f->lm[k], (size_t)LM_FIELDS * std::max(cap, 1) * sizeof(double)));
        CUDA_TRY(f, cudaMemsetAsync(f->lm[k], 0, (size_t)LM_FIELDS * std::max(cap, 1) * sizeof(double), f->stream));
        CUDA_TRY(f, cudaMalloc(&f->dids[k], std::max(cap, 1) * sizeof(int)));
    }
    f->haveSigMap = make_sigma_maps(f);
    CUDA_TRY(f, cudaMalloc(&f->d_xi0s, 23 * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&f->d_Xs[0], 23 * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&f->d_Xs[1], 23 * sizeof(double)));
    {
        // the deferred downdate tiles of the look-ahead correction run at the LOWEST priority and everything on the critical
        // path (f->stream, f->stream2) at the highest: the CTAs of the next chunk's factor kernel must be dispatched ahead
        // of the queued tile CTAs, not after them
        int prLeast = 0, prGreatest = 0;
        CUDA_TRY(f, cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest));
        f->prLeast = prLeast;
        f->prGreatest = prGreatest;
        CUDA_TRY(f, cudaStreamCreateWithPriority(&f->stream2, cudaStreamNonBlocking, prGreatest));
        CUDA_TRY(f, cudaStreamCreateWithPriority(&f->stream3, cudaStreamNonBlocking, prLeast));
        CUDA_TRY(f, cudaStreamCreateWithPriority(&f->stream4, cudaStreamNonBlocking, prGreatest));
        CUDA_TRY(f, cudaStreamCreateWithPriority(&f->stream5, cudaStreamNonBlocking, prGreatest));
    }
    CUDA_TRY(f, cudaEventCreateWithFlags(&f->evFork, cudaEventDisableTiming));
    CUDA_TRY(f, cudaEventCreateWithFlags(&f->evJoin, cudaEventDisableTiming));
    CUDA_TRY(f, cudaEventCreateWithFlags(&f->evG0, cudaEventDisableTiming));
    CUDA_TRY(f, cudaEventCreateWithFlags(&f->evGate, cudaEventDisableTiming));
    CUDA_TRY(f, cudaEventCreateWithFlags(&f->evStrip0, cudaEventDisableTiming));
    CUDA_TRY(f, cudaEventCreateWithFlags(&f->evStrip1, cudaEventDisableTiming));
    CUDA_TRY(f, cudaMalloc(&f->d_ctx, sizeof(RiccatiCtx)));
    {
        int rcf = alloc_frame(f, 64, std::max(cap, 1));
        if (rcf != EQVIO_OK) return rcf;
    }
    const size_t c1 = std::max(cap, 1);
    CUDA_TRY(f, cudaMalloc(&f->d_rows, c1 * ROWS_STRIDE * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&f->d_uv, c1 * UV_STRIDE * sizeof(double)));
    const size_t mMax = 2 * c1;
    const size_t ldzMax = (mMax + dimpMax + 1 + 7) & ~size_t(7);
    f->zElems = std::max(std::max(ldzMax * mMax, sigElems), (size_t)(f->ld / YB_T) * YB_TILE);
    CUDA_TRY(f, cudaMalloc(&f->d_Z, f->zElems * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&f->d_Ysplit, (size_t)(f->ld / TC_T) * TC_BLOCK_BYTES));
    CUDA_TRY(f, cudaMalloc(&f->d_Y2, (size_t)(f->ld / YB_T) * YB_TILE * sizeof(double)));
    {
        const size_t Tm = f->ld / YB_T + 1;  // level words | one tile counter per chunk (persistent deferred launches)
        CUDA_TRY(f, cudaMalloc(&f->d_lvl, (Tm * (Tm + 1) + c1 + 8) * sizeof(int)));
        CUDA_TRY(f, cudaMemsetAsync(f->d_lvl, 0, (Tm * (Tm + 1) + c1 + 8) * sizeof(int), f->stream));
    }
    CUDA_TRY(f, cudaMalloc(&f->d_Lout, ((size_t)dimpMax + mMax + NB) * NB * sizeof(double)));
    {
        // block sweep (mode 2): Z = [S; W^T] for at most BC_MAX_ROWS measurement rows, tile-blocked panels, one inverse per block
        const size_t mp = (std::min<size_t>(mMax, BC_MAX_ROWS) + BC_T - 1) / BC_T * BC_T;
        const size_t ldy = ((size_t)dimpMax + BC_T - 1) / BC_T * BC_T;
        const size_t zb = (mp + ldy) * mp, zp = 2 * (mp / BC_T + ldy / BC_T) * YB_TILE, mt = (mp / BC_T) * BC_LX;
        CUDA_TRY(f, cudaMalloc(&f->d_bcZ, zb * sizeof(double)));
        CUDA_TRY(f, cudaMalloc(&f->d_bcZp, zp * sizeof(double)));
        CUDA_TRY(f, cudaMalloc(&f->d_bcMt, mt * sizeof(double)));
        CUDA_TRY(f, cudaMemsetAsync(f->d_bcZp, 0, zp * sizeof(double), f->stream));  // the pad columns of the tiles travel with the bulk copies
        CUDA_TRY(f, cudaMemsetAsync(f->d_bcMt, 0, mt * sizeof(double), f->stream));
        CUDA_TRY(f, cudaMalloc(&f->d_bcCnt, (16 + 8 * (mp / BC_T) + 8) * sizeof(int)));  // counters [0, 16) | block-row flags
        CUDA_TRY(f, cudaMemsetAsync(f->d_bcCnt, 0, (16 + 8 * (mp / BC_T) + 8) * sizeof(int), f->stream));
    }
    CUDA_TRY(f, cudaMalloc(&f->d_Cblk, c1 * 6 * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&f->d_Gamma, (size_t)(dimpMax + 8) * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&f->d_Gamma2, (size_t)(dimpMax + 8) * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&f->d_ytilde, 2 * c1 * sizeof(double)));
    // everything the host reads back after an update lives in ONE device block with the layout of the pinned h_out:
    // gate scalars | gate flag (+ a constant 0) | status words | state estimate  -- a steady frame downloads it with one copy
    f->outOffSpec = ((3 * c1 * sizeof(double)) + 63) & ~size_t(63);
    f->outOffStatus = f->outOffSpec + 64;
    f->outOffEst = (f->outOffStatus + (1 + c1) * sizeof(int) + 63) & ~size_t(63);
    f->outBytes = (f->outOffEst + (23 + 3 * c1) * sizeof(double) + 63) & ~size_t(63);  // whole 16-byte words are copied
    CUDA_TRY(f, cudaMallocHost(&f->h_out, f->outBytes));
    if (cudaHostGetDevicePointer((void**)&f->hd_out, f->h_out, 0) != cudaSuccess) {
        cudaGetLastError();
        f->hd_out = nullptr;
    }
    CUDA_TRY(f, cudaMalloc(&f->d_outblk, f->outBytes));
    CUDA_TRY(f, cudaMemsetAsync(f->d_outblk, 0, f->outBytes, f->stream));
    f->d_gate = reinterpret_cast<double*>(f->d_outblk);
    f->d_spec = reinterpret_cast<int*>(f->d_outblk + f->outOffSpec);
    f->d_status = reinterpret_cast<int*>(f->d_outblk + f->outOffStatus);
    f->d_out = reinterpret_cast<double*>(f->d_outblk + f->outOffEst);
    CUDA_TRY(f, cudaMalloc(&f->d_normalM, 2 * 441 * sizeof(double)));
    if (!f->st.fastRiccati || f->st.coordinateChoice == EQVIO_COORD_NORMAL) {
        int rcs = alloc_sric(f);
        if (rcs != EQVIO_OK) return rcs;
    }
    CUDA_TRY(f, cudaMalloc(&f->d_keepI, c1 * sizeof(int)));
    // landmark-set changes arrive as ONE block: new positions | old-index map | new ids  (a single upload per change)
    f->mapBlkBytes = c1 * (3 * sizeof(double) + 2 * sizeof(int));
    CUDA_TRY(f, cudaMalloc(&f->d_mapblk, f->mapBlkBytes));
    f->d_newP = reinterpret_cast<double*>(f->d_mapblk);
    f->d_map = reinterpret_cast<int*>(f->d_mapblk + c1 * 3 * sizeof(double));
    f->d_newIds = f->d_map + c1;
    for (int i = 0; i < 2; ++i) CUDA_TRY(f, cudaEventCreate(&f->augEv[i]));
    for (int i = 0; i < 4; ++i) CUDA_TRY(f, cudaEventCreate(&f->stageEv[i]));
    return EQVIO_OK;
}

void group_identity_flat(double* g) {
    for (int i = 0; i < 23; ++i) g[i] = 0.0;
    g[6] = 1.0;
    g[16] = 1.0;
}

// Replace the whole device state: xi0 sensor, X = identity, landmarks (ids, p) with Q = identity,
// Sigma = diag(diagInternal).
int reset_state(eqvio_filter* f, const double sensor[23], int n, const int* ids, const double* p,
                const std::vector<double>& diag) {
    if (n > f->cap) {
        f->err = "landmark count exceeds the capacity the handle was created with";
        return EQVIO_ERR_CAPACITY;
    }
    std::memcpy(f->xi0s, sensor, 23 * sizeof(double));
    int rc;
    if ((rc = upload(f, f->d_xi0s, sensor, 23)) != EQVIO_OK) return rc;
    f->normalMValid = false;
    double gid[23];
    group_identity_flat(gid);
    if ((rc = upload(f, f->d_Xs[f->xcur], gid, 23)) != EQVIO_OK) return rc;
    f->ids.assign(ids, ids + n);
    if (n > 0) {
        std::vector<double> soa((size_t)LM_FIELDS * n);
        for (int i = 0; i < n; ++i) {
            soa[F_Q0X * n + i] = p[3 * i];
            soa[F_Q0Y * n + i] = p[3 * i + 1];
            soa[F_Q0Z * n + i] = p[3 * i + 2];
            soa[F_QW * n + i] = 1.0;
            soa[F_QX * n + i] = 0.0;
            soa[F_QY * n + i] = 0.0;
            soa[F_QZ * n + i] = 0.0;
            soa[F_QA * n + i] = 1.0;
        }
        double* h = static_cast<double*>(stage_alloc(f, soa.size() * sizeof(double)));
        if (!h) return EQVIO_ERR_CUDA;
        std::memcpy(h, soa.data(), soa.size() * sizeof(double));
        for (int fld = 0; fld < LM_FIELDS; ++fld)
            CUDA_TRY(f, cudaMemcpyAsync(f->lm[f->lmcur] + (size_t)fld * f->cap, h + (size_t)fld * n, n * sizeof(double),
                                        cudaMemcpyHostToDevice, f->stream));
        if ((rc = upload(f, f->dids[f->lmcur], ids, n)) != EQVIO_OK) return rc;
    }
    const int dimp = dimp_of(n);
    double* d_diag = f->d_Gamma;  // scratch, dimp <= dimpMax
    if ((rc = upload(f, d_diag, diag.data(), dimp)) != EQVIO_OK) return rc;
    dim3 grid(cdiv(dimp, 128), dimp);
    fill_diag_kernel<<<grid, 128, 0, f->stream>>>(f->Sig[f->cur], f->ld, dimp, d_diag);
    LAUNCH_CHECK(f, "fill_diag_kernel");
    CUDA_TRY(f, cudaStreamSynchronize(f->stream));
    return EQVIO_OK;
}

// Apply a landmark map (stable compaction + append) to lm / ids / Sigma.
// Device part of a landmark-set change: stable compaction / append of the landmark arrays and of Sigma (both ping-pong).
int launch_compaction(eqvio_filter* f, int newN, const int* d_map, const int* d_newIds, const double* d_newP, double newVar,
                      double newDepthVar) {
    if (newN > 0) {
        compact_landmarks_kernel<<<cdiv(newN, 128), 128, 0, f->stream>>>(f->lm[f->lmcur], f->lm[1 - f->lmcur], f->cap, f->dids[f->lmcur],
                                                                          f->dids[1 - f->lmcur], d_map, newN, d_newP, d_newIds,
                                                                          (const double*)f->d_Xs[f->xcur], f->d_Xs[1 - f->xcur]);
        LAUNCH_CHECK(f, "compact_landmarks_kernel");
        f->lmcur = 1 - f->lmcur;
        f->xcur = 1 - f->xcur;  // Sigma, landmarks and X flip together (see the kernel)
    }
    const int nb = 8 + newN;
    dim3 block(32, 8);
    dim3 grid(cdiv(nb, 32), cdiv(nb, 8));
    compact_sigma_kernel<<<grid, block, 0, f->stream>>>(f->Sig[f->cur], f->Sig[1 - f->cur], f->ld, d_map, newN, newVar, newDepthVar);
    LAUNCH_CHECK(f, "compact_sigma_kernel");
    f->cur = 1 - f->cur;
    return EQVIO_OK;
}

int apply_map(eqvio_filter* f, const std::vector<int>& map, const std::vector<int>& newIds, const std::vector<double>& newP,
              double newVar, double newDepthVar, bool newPOnDevice = false) {
    const int newN = (int)map.size();
    if (newN > f->cap) {
        f->err = "landmark count exceeds the capacity the handle was created with";
        return EQVIO_ERR_CAPACITY;
    }
    bool identity = (newN == (int)f->ids.size());
    for (int p = 0; identity && p < newN; ++p) identity = (map[p] == p);
    if (identity) return EQVIO_OK;
    int rc;
    std::vector<int> nids(newN);
    for (int p = 0; p < newN; ++p) nids[p] = map[p] >= 0 ? f->ids[map[p]] : newIds[-1 - map[p]];
    if (newN > 0) {
        {
            const size_t c1 = (size_t)std::max(f->cap, 1);
            const size_t offMap = c1 * 3 * sizeof(double), offIds = offMap + c1 * sizeof(int);
            // upload the prefix that is in use: positions (if any), the map, the new ids (if any)
            const size_t bytes = newIds.empty() ? offMap + (size_t)newN * sizeof(int) : offIds + newIds.size() * sizeof(int);
            unsigned char* h = static_cast<unsigned char*>(stage_alloc(f, bytes));
            if (!h) {
                f->err = "pinned staging allocation failed";
                return EQVIO_ERR_CUDA;
            }
            if (!newP.empty()) std::memcpy(h, newP.data(), newP.size() * sizeof(double));
            std::memcpy(h + offMap, map.data(), (size_t)newN * sizeof(int));
            if (!newIds.empty()) std::memcpy(h + offIds, newIds.data(), newIds.size() * sizeof(int));
            // without new landmarks (or with positions a kernel already wrote to d_newP) only the map and the ids travel
            const size_t from = (newIds.empty() || newPOnDevice) ? offMap : 0;
            CUDA_TRY(f, cudaMemcpyAsync(f->d_mapblk + from, h + from, bytes - from, cudaMemcpyHostToDevice, f->stream));
        }
    }
    if ((rc = launch_compaction(f, newN, f->d_map, f->d_newIds, f->d_newP, newVar, newDepthVar)) != EQVIO_OK) return rc;
    f->ids.swap(nids);
    return EQVIO_OK;
}

// removeOldLandmarks (VIOFilter.cpp:280-302) followed by an append of `addIds` with positions `addP`.
int remove_and_append(eqvio_filter* f, const std::vector<char>& keep, const std::vector<int>& addIds,
                      const std::vector<double>& addP, double newVar, double newDepthVar, bool addPOnDevice = false) {
    std::vector<int> map;
    map.reserve(f->ids.size() + addIds.size());
    for (int i = 0; i < (int)f->ids.size(); ++i)
        if (keep[i]) map.push_back(i);
    for (int k = 0; k < (int)addIds.size(); ++k) map.push_back(-1 - k);
    return apply_map(f, map, addIds, addP, newVar, newDepthVar, addPOnDevice);
}

// integrateUpToTime (VIOFilter.cpp:134-192), host part: segment lengths, time-weighted mean IMU, buffer pruning.
// *advanced = 0 reproduces the reference's `return false`.  Fills the frame header / IMU rows in h_frame.
int plan_integration(eqvio_filter* f, double newTime, int* advanced) {
    *advanced = 0;
    if (newTime <= f->time || f->time < 0 || f->buf.empty()) return EQVIO_OK;
    const eqvio_settings& s = f->st;
    // useDiscreteStateMatrix: stateMatrixADiscrete always differentiates liftVelocityDiscrete (EqFMatrices.cpp:24-41) whatever lift
    // the observer itself uses (VIOFilter.cpp:177), and it does so in the chart of the selected coordinate suite.
    const int n = (int)f->buf.size();
    if (n > f->maxSteps) {
        CUDA_TRY(f, cudaStreamSynchronize(f->stream));
        int rc = alloc_frame(f, 2 * n, f->yCap);
        if (rc != EQVIO_OK) return rc;
    }
    FrameHeader* hdr = reinterpret_cast<FrameHeader*>(f->h_frame);
    double* imu = reinterpret_cast<double*>(f->h_frame + f->offImu);
    double accT = 0.0, acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const double t0 = std::max(f->buf[i].stamp, f->time);
        const double t1 = i + 1 < n ? std::min(f->buf[i + 1].stamp, newTime) : newTime;
        const double dt = std::max(t1 - t0, 0.0);
        imu[13 * i] = dt;
        for (int k = 0; k < 12; ++k) imu[13 * i + 1 + k] = f->buf[i].v[k];
        accT += dt;
        for (int k = 0; k < 12; ++k) acc[k] = acc[k] + f->buf[i].v[k] * dt;
    }
    const double inv = 1.0 / accT;
    for (int k = 0; k < 12; ++k) hdr->fs.meanImu[k] = acc[k] * inv;
    hdr->fs.dtTotal = accT;
    hdr->fs.nsteps = n;
    hdr->fs.pad = 0;
    f->time = newTime;
    // prune, keeping the last sample with stamp < currentTime (VIOFilter.cpp:183-189)
    size_t k = 0;
    while (k < f->buf.size() && !(f->buf[k].stamp >= f->time)) ++k;
    if (k != 0) f->buf.erase(f->buf.begin(), f->buf.begin() + (k - 1));
    *advanced = 1;
    return EQVIO_OK;
}

// Device part of the propagation; the frame block must already be on its way to the device on f->stream.
// Two independent chains (VIOFilter.cpp:138 "the Riccati propagation ... does not affect the state propagation"):
//   stream : Riccati  -- context + Sigma_ss, landmark rows, strips, landmark-landmark block (reads X, Q *before*)
//   stream2: observer -- sensor part of every IMU segment, then the landmark part (writes the other X / lm buffers)
// Instantiations of the one-pass kernels by camera model / chart / lift form (kernels.cuh, gate_body): picked on the host per launch.
using MeasFn = decltype(&meas_kernel<-1, -1>);
using GateFn = decltype(&gate_kernel<-1, -1>);
using LiftFn = decltype(&lift_kernel<-1, -1>);
#define EQ_CAM_ROW(K, c) {K<CAM_PINHOLE, c>, K<CAM_RADTAN, c>, K<CAM_EQUIDISTANT, c>}
static MeasFn meas_fn(const eqvio_filter* f) {
    static const MeasFn t[3][3] = {EQ_CAM_ROW(meas_kernel, 0), EQ_CAM_ROW(meas_kernel, 1), EQ_CAM_ROW(meas_kernel, 2)};
    const int c = f->st.coordinateChoice, m = f->pend.cam.model;
    return (c >= 0 && c < 3 && m >= 0 && m < 3) ? t[c][m] : meas_kernel<-1, -1>;
}
static GateFn gate_fn(const eqvio_filter* f) {
    static const GateFn t[3][3] = {EQ_CAM_ROW(gate_kernel, 0), EQ_CAM_ROW(gate_kernel, 1), EQ_CAM_ROW(gate_kernel, 2)};
    const int c = f->st.coordinateChoice, m = f->pend.cam.model;
    return (c >= 0 && c < 3 && m >= 0 && m < 3) ? t[c][m] : gate_kernel<-1, -1>;
}
#undef EQ_CAM_ROW
using PrepFn = decltype(&riccati_prep_kernel<-1>);
static PrepFn prep_fn(const eqvio_filter* f) {
    static const PrepFn t[3] = {riccati_prep_kernel<0>, riccati_prep_kernel<1>, riccati_prep_kernel<2>};
    const int c = f->st.coordinateChoice;
    return (c >= 0 && c < 3) ? t[c] : riccati_prep_kernel<-1>;
}
static LiftFn lift_fn(const eqvio_filter* f) {
    static const LiftFn t[3][2] = {{lift_kernel<0, 0>, lift_kernel<0, 1>}, {lift_kernel<1, 0>, lift_kernel<1, 1>}, {lift_kernel<2, 0>, lift_kernel<2, 1>}};
    const int c = f->st.coordinateChoice;
    return (c >= 0 && c < 3) ? t[c][f->st.useDiscreteInnovationLift ? 1 : 0] : lift_kernel<-1, -1>;
}

void fill_prep_args(eqvio_filter* f, PrepArgs& a) {
    const eqvio_settings& s = f->st;
    a.xi0s = f->d_xi0s;
    a.Xs = f->d_Xs[f->xcur];
    a.XsOut = f->d_Xs[1 - f->xcur];
    a.ctx = f->d_ctx;
    a.steps = f->d_steps;
    a.fr = f->d_hdr;
    a.imu = f->d_imu;
    a.discreteLift = s.useDiscreteVelocityLift ? 1 : 0;
    a.qdiag[0] = s.velGyrNoise * s.velGyrNoise;
    a.qdiag[1] = s.velAccNoise * s.velAccNoise;
    a.qdiag[2] = s.velGyrBiasWalk * s.velGyrBiasWalk;
    a.qdiag[3] = s.velAccBiasWalk * s.velAccBiasWalk;
    a.pdiag[0] = s.biasOmegaProcessVariance;
    a.pdiag[1] = s.biasAccelProcessVariance;
    a.pdiag[2] = s.attitudeProcessVariance;
    a.pdiag[3] = s.positionProcessVariance;
    a.pdiag[4] = s.velocityProcessVariance;
    a.pdiag[5] = s.cameraAttitudeProcessVariance;
    a.pdiag[6] = s.cameraPositionProcessVariance;
    a.pdiag[7] = s.pointProcessVariance;
}
int enqueue_sric_step(eqvio_filter* f, const PrepArgs& a, int mode, const double* imuRow, int* clearFlag);

// fastRiccati: ONE Riccati step over the frame with the time-weighted mean IMU (VIOFilter.cpp:140-158) beside the observer integration of
// every buffered sample.  Euclid / InvDepth: the rank-27 structured kernels; Normal chart: the compact-block kernels (A = M A_euclid M^-1).
int enqueue_propagation(eqvio_filter* f) {
    const eqvio_settings& s = f->st;
    const int N = (int)f->ids.size();
    PrepArgs a;
    fill_prep_args(f, a);
    CUDA_TRY(f, cudaEventRecord(f->evFork, f->stream));
    CUDA_TRY(f, cudaStreamWaitEvent(f->stream2, f->evFork, 0));
    const int nsteps = reinterpret_cast<const FrameHeader*>(f->h_frame)->fs.nsteps;
    if (N > 0 && nsteps <= OBS_STAGE && f->fuseObserver) {
        observer_fused_kernel<<<cdiv(N, OBSF_LM), OBSF_THREADS, 0, f->stream2>>>(a, f->lm[f->lmcur], f->lm[1 - f->lmcur], f->dids[f->lmcur],
                                                                                 f->dids[1 - f->lmcur], f->cap, N, TL_SLOT(f));
        LAUNCH_CHECK(f, "observer_fused_kernel");
    } else {
        observer_sensor_kernel<<<1, 32, 0, f->stream2>>>(a, TL_SLOT(f));
        LAUNCH_CHECK(f, "observer_sensor_kernel");
        if (N > 0) {
            observer_landmark_kernel<<<cdiv(N, 64), 64, 0, f->stream2>>>(f->lm[f->lmcur], f->lm[1 - f->lmcur], f->dids[f->lmcur],
                                                                         f->dids[1 - f->lmcur], f->cap, N, f->d_steps, f->d_hdr, TL_SLOT(f));
            LAUNCH_CHECK(f, "observer_landmark_kernel");
        }
    }
    if (f->measHookNm > 0 && N > 0) {
        // C*, ytilde of the measured landmarks only read the landmarks the observer just integrated: right behind it on its stream,
        // beside the Riccati chain (the gate, which also needs the propagated Sigma, runs beside the sweep: enqueue_correction)
        const int nm = f->measHookNm;
        // (a plain launch: as a programmatic dependent of the observer it was measured 4.8 us SLOWER with a flushed L2 -- 7.9 vs 3.1 us)
        meas_fn(f)<<<cdiv(nm, 128), 128, 0, f->stream2>>>(f->lm[1 - f->lmcur], f->cap, f->d_lmOf, nm, f->d_y, f->d_hdr, s.coordinateChoice,
                                                           s.useEquivariantOutput ? 1 : 0, f->d_Cblk, f->d_ytilde, 1, 0, (const int*)(f->d_spec + 1),
                                                           f->d_yIdx, f->d_status, 1 + N, f->d_Gamma, dimp_of(N), (int*)nullptr, 0, TL_SLOT(f));
        LAUNCH_CHECK(f, "meas_kernel");
    }
    CUDA_TRY(f, cudaEventRecord(f->evJoin, f->stream2));
    if (s.coordinateChoice == EQVIO_COORD_NORMAL) {
        int rcs = enqueue_sric_step(f, a, 0, nullptr, f->d_spec);
        if (rcs != EQVIO_OK) return rcs;
    } else {
        const double* Sin = f->Sig[f->cur];
        double* Sout = f->Sig[1 - f->cur];
        // prologue + landmark rows in one launch; also re-arms the gate flag
        // (a programmatic dependent of the frame upload: its CTAs are resident when the upload's last store lands)
        launch_pdl(f, prep_fn(f), dim3(1 + cdiv(N, PREP_LM)), dim3(PREP_THREADS), (size_t)0, f->stream, a, Sin, Sout, f->ld, (double*)nullptr, f->d_spec,
                   (const double*)f->lm[f->lmcur], f->cap, N, (int)s.coordinateChoice, f->d_rows, TL_SLOT(f));
        LAUNCH_CHECK(f, "riccati_prep_kernel");
        if (N > 0) {
            const int nt = cdiv(N, TP);
            const bool fused = f->propFusion && !f->profiling;
            if (fused) {
                // the strip (sensor-landmark block only) beside the landmark-landmark kernel, which builds its own factors
                CUDA_TRY(f, cudaEventRecord(f->evStrip0, f->stream));
                CUDA_TRY(f, cudaStreamWaitEvent(f->stream5, f->evStrip0, 0));
                prop_strip_kernel<<<cdiv(N, PS_LM), PS_LM * PS_TPL, 0, f->stream5>>>(Sin, Sout, f->ld, N, f->d_ctx, f->d_rows, (double*)nullptr, TL_SLOT(f));
                LAUNCH_CHECK(f, "prop_strip_kernel");
                CUDA_TRY(f, cudaEventRecord(f->evStrip1, f->stream5));
            } else {
                launch_pdl(f, prop_strip_kernel, dim3(cdiv(N, PS_LM)), dim3(PS_LM * PS_TPL), (size_t)(0), f->stream, Sin, Sout, f->ld, N, f->d_ctx, f->d_rows, f->d_uv, TL_SLOT(f));
                LAUNCH_CHECK(f, "prop_strip_kernel");
            }
            int pk = prof_begin(f, PROF_PROP_LL);
            launch_pdl(f, prop_ll_kernel, dim3(dim3(nt, nt)), dim3(dim3(TP, TP)), (size_t)(0), f->stream, Sin, Sout, f->ld, N, f->d_ctx, f->d_rows, f->d_uv,
                       fused ? 1 : 0, TL_SLOT(f));
            prof_end(f, pk);
            LAUNCH_CHECK(f, "prop_ll_kernel");
            if (fused) CUDA_TRY(f, cudaStreamWaitEvent(f->stream, f->evStrip1, 0));
        }
        f->cur = 1 - f->cur;
    }
    CUDA_TRY(f, cudaStreamWaitEvent(f->stream, f->evJoin, 0));
    f->xcur = 1 - f->xcur;
    if (N > 0) f->lmcur = 1 - f->lmcur;
    return EQVIO_OK;
}

// Normal chart: M_s and its inverse depend on xi0's sensor part only -- computed once per xi0, outside any graph capture.
int ensure_normal_m(eqvio_filter* f) {
    if (f->normalMValid || f->capturing) return EQVIO_OK;
    sric::normal_m_sensor_kernel<<<1, 64, 0, f->stream>>>(f->d_xi0s, f->d_normalM, f->d_normalM + 441);
    LAUNCH_CHECK(f, "sric::normal_m_sensor_kernel");
    f->normalMValid = true;
    return EQVIO_OK;
}

int alloc_sric(eqvio_filter* f) {
    auto& w = f->sric;
    if (w.Ts) return EQVIO_OK;
    const size_t c1 = std::max(f->cap, 1);
    CUDA_TRY(f, cudaMalloc(&w.Ts, sric::SSIZE * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&w.Es, sric::SSIZE * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&w.Tl, c1 * sric::LSTRIDE * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&w.El, c1 * sric::LSTRIDE * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&w.uv, c1 * sric::GUV_STRIDE * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&w.ladder, (size_t)(sric::EXP_MAX_SQ + 1) * sric::SSIZE * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&w.dtBs, 252 * sizeof(double)));
    CUDA_TRY(f, cudaMalloc(&w.norm, sizeof(unsigned long long)));
    CUDA_TRY(f, cudaMalloc(&w.cc, sric::DA_SENSOR_EVALS * sizeof(SE3)));
    return EQVIO_OK;
}

// One Riccati step with the compact-block kernels (structured_riccati.cuh) on f->stream, Sig[cur] -> Sig[1 - cur]; the step's dt and mean
// IMU are a.fr's.  mode 0: E = I + dt A (one step over the frame: Normal chart with fastRiccati);  1: integrateRiccatiStateDiscrete
// (imuRow = the sample's row in d_imu);  2: integrateRiccatiStateAccurate.  No host synchronisation, nothing read from the host:
// capturable.  clearFlag: the first kernel of a steady update re-arms the gate flag.
int enqueue_sric_step(eqvio_filter* f, const PrepArgs& a, int mode, const double* imuRow, int* clearFlag) {
    const eqvio_settings& s = f->st;
    const int N = (int)f->ids.size();
    auto& w = f->sric;
    if (!w.Ts) {
        int rca = alloc_sric(f);
        if (rca != EQVIO_OK) return rca;
    }
    const bool normal = s.coordinateChoice == EQVIO_COORD_NORMAL;
    const double* Sin = f->Sig[f->cur];
    double* Sout = f->Sig[1 - f->cur];
    // sensor blocks of A, B for this step (the fused Sigma'_ss it also writes is overwritten by gprop_sensor_kernel below)
    prep_fn(f)<<<1 + cdiv(N, PREP_LM), PREP_THREADS, 0, f->stream>>>(a, Sin, Sout, f->ld, w.dtBs, clearFlag, f->lm[f->lmcur], f->cap, N,
                                                                               s.coordinateChoice, f->d_rows, TL_SLOT(f));
    LAUNCH_CHECK(f, "riccati_prep_kernel");
    // integrateRiccatiStateAccurate: T = dt [A B; 0 0], E = exp(T).  integrateRiccatiStateDiscrete (VIO_eqf.cpp:93-103,
    // useDiscreteStateMatrix): E = [A0tD, dt B] directly, with the numerically differentiated discrete state matrix
    // -- the products and the noise term below are then the same expressions (dt B (Q / dt) (dt B)^T = dt B Q B^T).
    const bool expo = mode == 2;
    double* Ts = expo ? w.Ts : w.Es;  // only the exponential maps T to a different E
    double* Tl = expo ? w.Tl : w.El;
    sric::fill_kernel<<<1 + cdiv(N, 128), 128, 0, f->stream>>>(f->d_ctx, w.dtBs, f->d_rows, N, Ts, Tl, w.norm);
    LAUNCH_CHECK(f, "sric::fill_kernel");
    if (normal) {
        int rcm = ensure_normal_m(f);  // a no-op inside a capture: vision_phase_a ran it before the capture began
        if (rcm != EQVIO_OK) return rcm;
        sric::normal_transform_kernel<<<1 + N, 128, 0, f->stream>>>(Ts, Tl, N, f->lm[f->lmcur], f->cap, f->d_normalM, f->d_normalM + 441);
        LAUNCH_CHECK(f, "sric::normal_transform_kernel");
    }
    if (mode == 0) {
        sric::add_identity_kernel<<<cdiv(std::max(3 * N, SENSOR_DIM), 128), 128, 0, f->stream>>>(w.Es, w.El, N);
        LAUNCH_CHECK(f, "sric::add_identity_kernel");
    } else if (mode == 1) {
        sric::discrete_a_sensor_kernel<<<1, 64, 0, f->stream>>>(s.coordinateChoice, f->d_xi0s, a.Xs, imuRow, w.Es, w.cc);
        LAUNCH_CHECK(f, "sric::discrete_a_sensor_kernel");
        if (N > 0) {
            sric::discrete_a_landmark_kernel<<<N, 64, 0, f->stream>>>(f->lm[f->lmcur], f->cap, N, s.coordinateChoice, imuRow, w.cc, w.El);
            LAUNCH_CHECK(f, "sric::discrete_a_landmark_kernel");
        }
    } else {
        sric::exp_norm_kernel<<<cdiv(SENSOR_DIM + 3 * N, 128), 128, 0, f->stream>>>(w.Ts, w.Tl, N, w.norm);
        LAUNCH_CHECK(f, "sric::exp_norm_kernel");
        sric::exp_sensor_kernel<<<1, sric::EXPS_THREADS, 0, f->stream>>>(w.Ts, w.norm, w.ladder, w.Es);
        LAUNCH_CHECK(f, "sric::exp_sensor_kernel");
        if (N > 0) {
            sric::exp_landmark_kernel<<<N, 128, 0, f->stream>>>(w.Ts, w.Tl, w.norm, w.ladder, w.El, N);
            LAUNCH_CHECK(f, "sric::exp_landmark_kernel");
        }
    }
    sric::NoiseArgs nz;
    for (int k = 0; k < 4; ++k) nz.q[k] = a.qdiag[k];
    for (int k = 0; k < 8; ++k) nz.p[k] = a.pdiag[k];
    sric::gprop_sensor_kernel<<<1, 448, 0, f->stream>>>(w.Es, nz, f->d_ctx, Sin, Sout, f->ld);
    LAUNCH_CHECK(f, "sric::gprop_sensor_kernel");
    if (N > 0) {
        sric::gprop_strip_kernel<<<cdiv(N, sric::GS_LM), sric::GS_LM * 32, 0, f->stream>>>(w.Es, w.El, nz, f->d_ctx, Sin, Sout, f->ld, N, w.uv);
        LAUNCH_CHECK(f, "sric::gprop_strip_kernel");
        const int nt = cdiv(N, sric::GTP);
        int pk = prof_begin(f, PROF_PROP_LL);
        sric::gprop_ll_kernel<<<dim3(nt, nt), dim3(sric::GTP, sric::GTP), sric::GLL_SMEM, f->stream>>>(w.El, w.uv, f->d_ctx, Sin, Sout, f->ld, N);
        prof_end(f, pk);
        LAUNCH_CHECK(f, "sric::gprop_ll_kernel");
    }
    f->cur = 1 - f->cur;
    return EQVIO_OK;
}

// fastRiccati = false: per buffered IMU sample, integrateRiccatiStateAccurate (or ...Discrete with useDiscreteStateMatrix) followed by
// that sample's integrateObserverState -- VIOFilter.cpp:160-178.  Block-structured kernels of structured_riccati.cuh: no dense
// matrix, no library call, no host synchronisation.
int enqueue_propagation_per_sample(eqvio_filter* f) {
    const eqvio_settings& s = f->st;
    const int N = (int)f->ids.size();
    const FrameHeader* hh = reinterpret_cast<const FrameHeader*>(f->h_frame);
    const double* himu = reinterpret_cast<const double*>(f->h_frame + f->offImu);
    const int nsteps = hh->fs.nsteps;
    int rc;
    std::vector<FrameHeader> hdrs(nsteps);
    for (int i = 0; i < nsteps; ++i) {
        hdrs[i] = *hh;
        for (int k = 0; k < 12; ++k) hdrs[i].fs.meanImu[k] = himu[13 * i + 1 + k];
        hdrs[i].fs.dtTotal = himu[13 * i];
        hdrs[i].fs.nsteps = 1;
    }
    if ((rc = upload(f, f->d_hdrSteps, hdrs.data(), (size_t)nsteps)) != EQVIO_OK) return rc;
    for (int i = 0; i < nsteps; ++i) {
        const double dt = himu[13 * i];
        PrepArgs a;
        fill_prep_args(f, a);
        a.steps = f->d_steps + i;
        a.fr = f->d_hdrSteps + i;
        a.imu = f->d_imu + (size_t)13 * i;
        if (dt > 0 && (rc = enqueue_sric_step(f, a, s.useDiscreteStateMatrix ? 1 : 2, f->d_imu + (size_t)13 * i, nullptr)) != EQVIO_OK) return rc;
        observer_sensor_kernel<<<1, 32, 0, f->stream>>>(a, TL_SLOT(f));
        LAUNCH_CHECK(f, "observer_sensor_kernel");
        if (N > 0) {
            observer_landmark_kernel<<<cdiv(N, 64), 64, 0, f->stream>>>(f->lm[f->lmcur], f->lm[1 - f->lmcur], f->dids[f->lmcur],
                                                                        f->dids[1 - f->lmcur], f->cap, N, f->d_steps + i, f->d_hdrSteps + i, TL_SLOT(f));
            LAUNCH_CHECK(f, "observer_landmark_kernel");
            f->lmcur = 1 - f->lmcur;
        }
        f->xcur = 1 - f->xcur;
    }
    return EQVIO_OK;
}

int enqueue_gate(eqvio_filter* f, int N, bool clearFlag) {
    if (clearFlag) CUDA_TRY(f, cudaMemsetAsync(f->d_spec, 0, sizeof(int), f->stream));
    launch_pdl(f, gate_fn(f), dim3(cdiv(N, 128)), dim3(128), (size_t)(0), f->stream, f->lm[f->lmcur], f->cap, N, f->Sig[f->cur], f->ld, f->d_measIdx, f->d_y,
                                                      f->d_hdr, f->st.coordinateChoice, f->d_gate, f->st.outlierThresholdAbs,
                                                      f->st.outlierThresholdProb, f->d_spec, TL_SLOT(f));
    LAUNCH_CHECK(f, "gate_kernel");
    return EQVIO_OK;
}

int enqueue_correction(eqvio_filter* f, int nm, const int* guard, bool fuseGate = false, bool fuseEst = false, bool lateGate = false);
struct FramePlan;

// The whole device side of a steady frame (no landmark enters or leaves before the gate): frame upload,
// propagation, gate, guarded correction, result downloads into the fixed pinned block.  Issued either directly
// or under stream capture (then replayed as one CUDA graph).
// plan: landmark-set change decided from the ids alone (lost ids pruned, new ids appended); null for a frame without one.
struct FramePlan {
    int Nnew = 0, nNew = 0;
    std::vector<int> nids;  // ids of the new state
};
int enqueue_steady_update(eqvio_filter* f, int N, int nm, const FramePlan* plan) {
    int rc;
    const eqvio_settings& s = f->st;
    const bool zc = f->zeroCopy && f->hd_frame && f->hd_out;
    if (zc) {
        // + prefetch of the covariance (the rows in use) into L2 beside the copy
        const size_t pfBytes = f->prefetchSigma ? (size_t)f->ld * dimp_of(N) * sizeof(double) : 0;
        PrefetchList small = {};
        if (f->prefetchSigma) {
            const size_t c1 = std::max(f->cap, 1);
            auto add = [&](const void* p, size_t bytes, unsigned stride) {
                if (p && bytes && small.n < PF_MAX) {
                    small.p[small.n] = reinterpret_cast<const char*>(p);
                    small.bytes[small.n] = (unsigned)std::min(bytes, (size_t)0xFFFFFFF0u);
                    small.stride[small.n++] = stride;
                }
            };
            // read first by the prologue / the observer chain: every line
            add(f->d_xi0s, 23 * sizeof(double), 128);
            add(f->d_Xs[f->xcur], 23 * sizeof(double), 128);
            add(f->lm[f->lmcur], LM_FIELDS * c1 * sizeof(double), 128);
            add(f->dids[f->lmcur], c1 * sizeof(int), 128);
            // (touching every scratch / output buffer as well -- in case the first access to their pages paid a cold translation -- was
            // measured and changes nothing: the cold-L2 penalty of the one-pass kernels, +8 us in the prologue, +4 us in the measurement
            // rows, is not a data miss)
        }
        block_copy_kernel<<<2 + (pfBytes ? 64 : 0), 256, 0, f->stream>>>(reinterpret_cast<const double2*>(f->hd_frame), reinterpret_cast<double2*>(f->d_frame),
                                                                        (int)(f->frameBytes / 16), 2, reinterpret_cast<const char*>(f->Sig[f->cur]), pfBytes,
                                                                        small, TL_SLOT(f));
        LAUNCH_CHECK(f, "block_copy_kernel<frame>");
    } else {
        CUDA_TRY(f, cudaMemcpyAsync(f->d_frame, f->h_frame, f->frameBytes, cudaMemcpyHostToDevice, f->stream));
    }
    const bool fuseEst = f->fuseSmall && f->corrMode != 1;
    const bool fuseGate = fuseEst && !plan;  // with a landmark-set change the gate sees the OLD state, the rows the NEW one
    // block sweep without a landmark-set change: rows early (observer stream), gate late (beside the sweep)
    const bool lateGate = fuseGate && f->earlyRows && f->corrMode == 2 && 2 * nm <= BC_MAX_ROWS && !f->downdateTC && !f->profiling && nm > 0;
    f->measHookNm = lateGate ? nm : 0;
    rc = enqueue_propagation(f);
    f->measHookNm = 0;
    if (rc != EQVIO_OK) return rc;
    f->steadySplit = !f->capturing;  // stage brackets inside the update exist only with plain launches (a replayed graph is one bracket)
    if (f->steadySplit) stage_mark(f, 1);
    if (!fuseGate && (rc = enqueue_gate(f, N, false)) != EQVIO_OK) return rc;  // riccati_prep_kernel re-armed the flag
    int Nout = N;
    if (plan) {
        // everything below reads its per-frame data (keep flags, new measurement indices, map, new ids, count) from the frame
        // block, so the same graph serves every frame with these sizes
        if (plan->nNew > 0) {
            launch_pdl(f, new_landmark_kernel, dim3(1), dim3(256), (size_t)0, f->stream, (const double*)f->d_gate, (const int*)f->d_keepF, N,
                       s.useMedianDepth ? 1 : 0, s.initialSceneDepth, (const FrameHeader*)f->d_hdr, (const double*)f->d_y,
                       (const int*)f->d_newMeasF, 0, (const int*)f->d_counts, f->d_newP);
            LAUNCH_CHECK(f, "new_landmark_kernel");
        }
        if ((rc = launch_compaction(f, plan->Nnew, f->d_mapF, f->d_newIdsF, f->d_newP, s.initialPointVariance, -1.0)) != EQVIO_OK) return rc;
        f->ids = plan->nids;
        Nout = plan->Nnew;
    }
    if (f->steadySplit) stage_mark(f, 2);
    // maxOutliers == 0 (featureRetention = 1, or (1 - featureRetention) n < 1): removeOutliers removes nothing whatever the gate
    // says (VIOFilter.cpp:304-364), so the correction must not be guarded by the gate flag -- d_spec + 1 is a constant 0
    if ((rc = enqueue_correction(f, nm, f->pend.ignoreGate ? f->d_spec + 1 : f->d_spec, fuseGate, fuseEst, lateGate)) != EQVIO_OK) return rc;
    // stateEstimate() is what every caller asks for next (main_opt.cpp:225, main_sim.cpp:146): produced here, by the lift itself
    // in the fused form
    if (!fuseEst) {
        launch_pdl(f, state_estimate_kernel, dim3(cdiv(Nout, 128)), dim3(128), (size_t)(0), f->stream, f->lm[f->lmcur], f->cap, Nout, f->d_xi0s,
                   f->d_Xs[f->xcur], f->d_out, TL_SLOT(f));
        LAUNCH_CHECK(f, "state_estimate_kernel");
    }
    // one download: gate scalars, gate flag, status words, state estimate (the block has the layout of h_out)
    const size_t outBytes = f->outOffEst + (23 + 3 * (size_t)Nout) * sizeof(double);
    if (zc) {
        launch_pdl(f, block_copy_kernel, dim3(2), dim3(256), (size_t)0, f->stream, reinterpret_cast<const double2*>(f->d_outblk),
                   reinterpret_cast<double2*>(f->hd_out), (int)((outBytes + 15) / 16), 2, (const char*)nullptr, (size_t)0, PrefetchList{}, TL_SLOT(f));
        LAUNCH_CHECK(f, "block_copy_kernel<result>");
    } else {
        CUDA_TRY(f, cudaMemcpyAsync(f->h_out, f->d_outblk, outBytes, cudaMemcpyDeviceToHost, f->stream));
    }
    return EQVIO_OK;
}

// ---- process_vision, phase A: plan, classify the frame, enqueue ---------------------------------------------
int vision_phase_a(eqvio_filter* f, double stamp, int n, const int* ids, const double* y, const eqvio_camera* cam) {
    auto& P = f->pend;
    P = eqvio_filter::Pending();
    stage_reset(f);
    f->lastOutliers.clear();
    f->estValid = false;
    if (n < 0 || (n > 0 && (!ids || !y)) || !cam) {
        f->err = "invalid measurement arguments";
        return EQVIO_ERR_INVALID_ARG;
    }
    for (int j = 1; j < n; ++j)
        if (ids[j] <= ids[j - 1]) {
            f->err = "measurement ids must be strictly ascending";
            return EQVIO_ERR_INVALID_ARG;
        }
    if (cam->model != EQVIO_CAMERA_PINHOLE && cam->model != EQVIO_CAMERA_RADTAN && cam->model != EQVIO_CAMERA_EQUIDISTANT) {
        f->err = "unsupported camera model";
        return EQVIO_ERR_UNSUPPORTED;
    }
    int rc;
    if (n > f->yCap) {  // a measurement may list more ids than the state can hold (gated later)
        CUDA_TRY(f, cudaStreamSynchronize(f->stream));
        if ((rc = alloc_frame(f, f->maxSteps, 2 * n)) != EQVIO_OK) return rc;
    }
    int advanced = 0;
    if ((rc = plan_integration(f, stamp, &advanced)) != EQVIO_OK) return rc;
    if (!advanced || !f->initialised) return EQVIO_OK;  // VIOFilter.cpp:198-199
    P.active = true;
    P.n = n;
    P.mids.assign(ids, ids + n);
    P.my.assign(y, y + 2 * n);
    P.cam = to_camera(cam);
    const int N = (int)f->ids.size();
    // frame block: camera, pixels, per-landmark measurement index
    FrameHeader* hdr = reinterpret_cast<FrameHeader*>(f->h_frame);
    hdr->cam = P.cam;
    if (n > 0) std::memcpy(f->h_frame + f->offY, y, 2 * (size_t)n * sizeof(double));
    int* hMeasIdx = reinterpret_cast<int*>(f->h_frame + f->offMeasIdx);
    int* hLmOf = reinterpret_cast<int*>(f->h_frame + f->offLmOf);
    int* hYIdx = reinterpret_cast<int*>(f->h_frame + f->offYIdx);
    P.measIdx.assign(N, -1);
    P.keep.assign(N, 1);
    bool anyLost = false;
    int matched = 0;
    for (int i = 0; i < N; ++i) {
        // the measurement ids are strictly ascending: binary search instead of a hash map
        const int* it = std::lower_bound(ids, ids + n, f->ids[i]);
        if (it != ids + n && *it == f->ids[i]) {
            const int j = (int)(it - ids);
            P.measIdx[i] = j;
            // rows of the correction follow the STATE order (chunks then cover contiguous rows / columns of Sigma);
            // the update does not depend on the row order (VIO_eqf.cpp:116-131 with R = sigma^2 I)
            hLmOf[matched] = i;
            hYIdx[matched] = j;
            ++matched;
        } else if (f->st.removeLostLandmarks) {
            P.keep[i] = 0;  // removeOldLandmarks, VIOFilter.cpp:203-205
            anyLost = true;
        }
        hMeasIdx[i] = P.measIdx[i];
    }
    P.oldIds = f->ids;
    f->h_lmOfSorted.assign(hLmOf, hLmOf + matched);
    const bool anyNew = matched < n;
    const size_t maxOutliers = (size_t)((1.0 - f->st.featureRetention) * n);
    // Planned frame: every decision that shapes the launch sequence is known now.  That is the case when nothing enters or
    // leaves before the gate, and also when ids are lost (pruned, VIOFilter.cpp:203-205) or new (appended with bearing x median
    // depth, :258-278) -- those sets follow from the ids alone, and the new landmarks' positions are computed on the device under
    // the same "no gate trips" assumption the speculative correction makes (exact redo in phase C otherwise).
    const bool densePath = !f->st.fastRiccati;  // per-sample Riccati variants: per-kernel launches (the Normal chart with fastRiccati is a steady path)
    int nLost = 0;
    for (int i = 0; i < N; ++i) nLost += P.keep[i] ? 0 : 1;
    const int nNewIds = n - matched;
    const bool changeOk = (!anyNew && !anyLost) || (f->specNew && (N - nLost) + nNewIds <= f->cap && (N - nLost) + nNewIds > 0);
    P.steady = f->speculate && f->corrMode != 1 && N > 0 && n > 0 && !densePath && changeOk;
    P.ignoreGate = maxOutliers == 0;
    if (P.steady) {
        FramePlan plan;
        const bool change = anyNew || anyLost;
        int matchedNew = matched;
        if (change) {
            int* hKeep = reinterpret_cast<int*>(f->h_frame + f->offKeepF);
            int* hNewMeas = reinterpret_cast<int*>(f->h_frame + f->offNewMeasF);
            int* hMap = reinterpret_cast<int*>(f->h_frame + f->offMapF);
            int* hNewIds = reinterpret_cast<int*>(f->h_frame + f->offNewIdsF);
            int* hCounts = reinterpret_cast<int*>(f->h_frame + f->offCounts);
            // new state = kept landmarks in their order, then the new ids in measurement (ascending id) order; its rows follow
            int p = 0, rows = 0;
            for (int i = 0; i < N; ++i) {
                hKeep[i] = P.keep[i] ? 1 : 0;
                if (!P.keep[i]) continue;
                hMap[p] = i;
                plan.nids.push_back(f->ids[i]);
                if (P.measIdx[i] >= 0) {
                    hLmOf[rows] = p;
                    hYIdx[rows] = P.measIdx[i];
                    ++rows;
                }
                ++p;
            }
            std::vector<char> inState(n, 0);
            for (int i = 0; i < N; ++i)
                if (P.measIdx[i] >= 0) inState[P.measIdx[i]] = 1;
            int k = 0;
            for (int j = 0; j < n; ++j)
                if (!inState[j]) {
                    hNewMeas[k] = j;
                    hNewIds[k] = ids[j];
                    hMap[p] = -1 - k;
                    plan.nids.push_back(ids[j]);
                    hLmOf[rows] = p;
                    hYIdx[rows] = j;
                    ++rows;
                    ++p;
                    ++k;
                }
            hCounts[0] = k;
            plan.Nnew = p;
            plan.nNew = k;
            matchedNew = rows;
            P.specNewCount = k;
            f->h_lmOfSorted.assign(hLmOf, hLmOf + rows);
        }
        const int Nout = change ? plan.Nnew : N;
        const int nm = change ? matchedNew : n;
        P.speculated = true;
        P.gated = true;
        P.measKept.assign(n, 1);
        P.h_gate = reinterpret_cast<double*>(f->h_out);
        P.h_spec = reinterpret_cast<int*>(f->h_out + f->outOffSpec);
        P.h_status = reinterpret_cast<int*>(f->h_out + f->outOffStatus);
        P.nStatus = 1 + Nout;
        if (f->st.coordinateChoice == EQVIO_COORD_NORMAL && (rc = ensure_normal_m(f)) != EQVIO_OK) return rc;
        stage_mark(f, 0);
        // kernel arguments derived from the row -> landmark map are baked into a captured graph: replay only when that map
        // is the identity (every landmark of the updated state measured)
        const bool graphOk = f->useGraph && !f->profiling && nm == Nout;
        if (!graphOk) {
            if ((rc = enqueue_steady_update(f, N, nm, change ? &plan : nullptr)) != EQVIO_OK) return rc;
            stage_mark(f, 3);
        } else {
            const eqvio_settings& st = f->st;
            std::vector<int> key = {N, n, f->cur, f->lmcur, f->xcur, f->chunkLm, f->corrMode, st.coordinateChoice, st.useDiscreteVelocityLift,
                                    st.useDiscreteInnovationLift, st.useEquivariantOutput, f->maxSteps, f->yCap,
                                    change ? 1 : 0, Nout, plan.nNew > 0 ? 1 : 0, P.ignoreGate ? 1 : 0, P.cam.model,
                                    // the observer kernel form is chosen on the host from the number of buffered IMU segments
                                    reinterpret_cast<const FrameHeader*>(f->h_frame)->fs.nsteps <= OBS_STAGE ? 1 : 0};
            auto it = f->graphs.find(key);
            if (it == f->graphs.end()) {
                if (f->graphs.size() >= 32) {  // evict the least recently used
                    auto victim = f->graphs.begin();
                    for (auto g = f->graphs.begin(); g != f->graphs.end(); ++g)
                        if (g->second.lastUse < victim->second.lastUse) victim = g;
                    cudaGraphExecDestroy(victim->second.exec);
                    f->graphs.erase(victim);
                }
                // capture for the ping-pong indices (c, l, x); nothing executes, the indices are restored afterwards
                auto capture = [&](int c, int l, int x, eqvio_filter::GraphEntry& ge) -> int {
                    const int c0 = f->cur, l0 = f->lmcur, x0 = f->xcur;
                    const long long launches0 = f->launches;
                    f->cur = c;
                    f->lmcur = l;
                    f->xcur = x;
                    cudaGraph_t graph = nullptr;
                    cudaError_t cb = cudaStreamBeginCapture(f->stream, cudaStreamCaptureModeThreadLocal);
                    if (cb != cudaSuccess) {
                        f->cur = c0; f->lmcur = l0; f->xcur = x0;
                        f->err = std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(cb);
                        return EQVIO_ERR_CUDA;
                    }
                    std::vector<int> ids0 = f->ids;  // a captured landmark-set change assigns f->ids: the replay below does it for real
                    f->capturing = true;
                    int rcc = enqueue_steady_update(f, N, nm, change ? &plan : nullptr);
                    f->capturing = false;
                    cudaError_t ce = cudaStreamEndCapture(f->stream, &graph);
                    f->ids.swap(ids0);
                    ge.cur2 = f->cur;
                    ge.lmcur2 = f->lmcur;
                    ge.xcur2 = f->xcur;
                    ge.launches = f->launches - launches0;
                    f->cur = c0;  // capture only records: the state flips when the graph is launched below
                    f->lmcur = l0;
                    f->xcur = x0;
                    f->launches = launches0;
                    if (rcc != EQVIO_OK) return rcc;
                    if (ce != cudaSuccess || !graph) {
                        f->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce);
                        return EQVIO_ERR_CUDA;
                    }
                    ce = cudaGraphInstantiate(&ge.exec, graph, 0);
                    cudaGraphDestroy(graph);
                    if (ce != cudaSuccess) {
                        f->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce);
                        return EQVIO_ERR_CUDA;
                    }
                    ++f->graphCaptures;
                    return EQVIO_OK;
                };
                eqvio_filter::GraphEntry ge;
                if ((rc = capture(f->cur, f->lmcur, f->xcur, ge)) != EQVIO_OK) return rc;
                it = f->graphs.emplace(key, ge).first;
                // ... and its twin with every index flipped: a frame that flips the buffers an odd number of times (a landmark-set change
                // in augmentLandmarkStates) would otherwise meet this shape again as a NEW key later, ~3 ms of capture + instantiation in
                // the middle of a run instead of beside the first one
                if (f->graphs.size() < 32) {
                    std::vector<int> twin = key;
                    twin[2] = 1 - f->cur;
                    twin[3] = 1 - f->lmcur;
                    twin[4] = 1 - f->xcur;
                    if (f->graphs.find(twin) == f->graphs.end()) {
                        eqvio_filter::GraphEntry gt;
                        if ((rc = capture(1 - f->cur, 1 - f->lmcur, 1 - f->xcur, gt)) != EQVIO_OK) return rc;
                        f->graphs.emplace(twin, gt);
                        it = f->graphs.find(key);
                    }
                }
            }
            it->second.lastUse = ++f->graphClock;
            CUDA_TRY(f, cudaGraphLaunch(it->second.exec, f->stream));
            stage_mark(f, 3);
            f->cur = it->second.cur2;
            f->lmcur = it->second.lmcur2;
            f->xcur = it->second.xcur2;
            f->launches += it->second.launches;
            ++f->graphLaunches;
            if (change) f->ids = plan.nids;
        }
        P.corrected = true;
        return EQVIO_OK;
    }
    // general frame: upload, propagate, gate; decisions follow in phase B
    CUDA_TRY(f, cudaMemcpyAsync(f->d_frame, f->h_frame, f->frameBytes, cudaMemcpyHostToDevice, f->stream));
    stage_mark(f, 0);
    if ((rc = densePath ? enqueue_propagation_per_sample(f) : enqueue_propagation(f)) != EQVIO_OK) return rc;
    stage_mark(f, 1);
    if (N > 0) {
        if ((rc = enqueue_gate(f, N, true)) != EQVIO_OK) return rc;
        if ((rc = download_async(f, &P.h_gate, f->d_gate, 3 * (size_t)N)) != EQVIO_OK) return rc;
        if ((rc = download_async(f, &P.h_spec, f->d_spec, 1)) != EQVIO_OK) return rc;
        P.gated = true;
    }
    return EQVIO_OK;
}

// ---- bookkeeping decisions on the host (indices refer to P.oldIds, the state when the gate ran) -----------
// removeOutliers (VIOFilter.cpp:304-364) from the gate scalars; useGate = false assumes that no landmark is gated.
void decide_outliers(eqvio_filter* f, bool useGate, std::vector<char>& outlier) {
    auto& P = f->pend;
    const eqvio_settings& s = f->st;
    const int N = (int)P.oldIds.size();
    const int n = P.n;
    outlier.assign(N, 0);
    P.measKept.assign(n, 1);
    f->lastOutliers.clear();
    if (!useGate || !P.h_gate) return;
    const double* errAbs = P.h_gate;
    const double* errProb = P.h_gate + N;
    const size_t maxOutliers = (size_t)((1.0 - s.featureRetention) * n);
    std::vector<int> order;  // measured, kept landmarks in ascending id like the reference's std::map iteration
    for (int i = 0; i < N; ++i)
        if (P.keep[i] && P.measIdx[i] >= 0) order.push_back(i);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return P.oldIds[a] < P.oldIds[b]; });
    std::vector<int> proposed;
    std::map<int, double> absOut, probOut;
    for (int i : order)
        if (errAbs[i] > s.outlierThresholdAbs) {
            absOut[i] = errAbs[i];
            proposed.push_back(i);
        }
    for (int i : order) {
        if (absOut.count(i)) continue;
        if (errProb[i] > s.outlierThresholdProb) {
            probOut[i] = errProb[i];
            proposed.push_back(i);
        }
    }
    std::sort(proposed.begin(), proposed.end(), [&](int a, int b) {
        if (absOut.count(a)) {
            if (absOut.count(b)) return absOut.at(a) < absOut.at(b);
            return false;
        }
        if (absOut.count(b)) return true;
        return probOut.at(a) < probOut.at(b);
    });
    std::reverse(proposed.begin(), proposed.end());
    if (proposed.size() > maxOutliers) proposed.resize(maxOutliers);
    for (int i : proposed) {
        outlier[i] = 1;
        P.measKept[P.measIdx[i]] = 0;
        f->lastOutliers.push_back(P.oldIds[i]);
    }
}

int launch_correction(eqvio_filter* f, const int* guard);

// ---- phase B: compaction and correction launches ---------------------------------------------------------
int vision_phase_b(eqvio_filter* f) {
    auto& P = f->pend;
    if (!P.active || P.steady) return EQVIO_OK;  // a steady frame was enqueued in full by phase A
    const eqvio_settings& s = f->st;
    int rc;
    const int N = (int)f->ids.size();
    const int n = P.n;
    std::vector<char> measInState(n, 0);
    for (int i = 0; i < N; ++i)
        if (P.measIdx[i] >= 0) measInState[P.measIdx[i]] = 1;
    bool anyNew = false;
    for (int j = 0; j < n; ++j) anyNew |= !measInState[j];
    const size_t maxOutliers = (size_t)((1.0 - s.featureRetention) * n);
    // Speculation: with no new ids the only thing the host needs from the device is "did any landmark trip a
    // gate".  The correction is launched right away, guarded on the device by that flag; phase C redoes it
    // through the exact path in the (rare) case the flag came back set.
    // Frames that bring NEW ids speculate too: the positions of the new landmarks (bearing x median scene depth) are computed
    // on the device from the gate kernel's depths under the same "no gate trips" assumption, so nothing waits for the host.
    int nNewIds = 0, nKeptOld = 0;
    for (int j = 0; j < n; ++j) nNewIds += !measInState[j];
    for (int i = 0; i < N; ++i) nKeptOld += P.keep[i] ? 1 : 0;
    const bool specNew = f->speculate && f->specNew && f->corrMode != 1 && P.gated && anyNew && maxOutliers > 0 && N > 0 &&
                         nKeptOld + nNewIds <= f->cap;  // (over capacity: the exact path reports the error)
    P.speculated = f->speculate && f->corrMode != 1 && P.gated && (!anyNew || specNew) && maxOutliers > 0;
    const bool noGateNeeded = !P.gated || (maxOutliers == 0 && !(anyNew && s.useMedianDepth));
    if (!P.speculated && !noGateNeeded) CUDA_TRY(f, cudaStreamSynchronize(f->stream));
    std::vector<char> outlier;
    decide_outliers(f, !P.speculated && !noGateNeeded, outlier);
    std::vector<char> keep(N);
    for (int i = 0; i < N; ++i) keep[i] = P.keep[i] && !outlier[i];

    // addNewLandmarks (VIOFilter.cpp:258-278)
    std::vector<int> addIds;
    std::vector<double> addP;
    if (anyNew && specNew) {
        std::vector<int> keepI(N), newMeas;
        for (int i = 0; i < N; ++i) keepI[i] = keep[i] ? 1 : 0;
        for (int j = 0; j < n; ++j)
            if (!measInState[j]) {
                addIds.push_back(P.mids[j]);
                newMeas.push_back(j);
            }
        P.specNewCount = (int)addIds.size();
        if ((int)newMeas.size() > f->newMeasCap) {
            CUDA_TRY(f, cudaStreamSynchronize(f->stream));
            cudaFree(f->d_newMeas);
            f->newMeasCap = 2 * (int)newMeas.size() + 64;
            CUDA_TRY(f, cudaMalloc(&f->d_newMeas, f->newMeasCap * sizeof(int)));
        }
        if ((rc = upload(f, f->d_keepI, keepI.data(), (size_t)N)) != EQVIO_OK) return rc;
        if ((rc = upload(f, f->d_newMeas, newMeas.data(), newMeas.size())) != EQVIO_OK) return rc;
        launch_pdl(f, new_landmark_kernel, dim3(1), dim3(256), (size_t)0, f->stream, (const double*)f->d_gate, (const int*)f->d_keepI, N,
                   s.useMedianDepth ? 1 : 0, s.initialSceneDepth, (const FrameHeader*)f->d_hdr, (const double*)f->d_y, (const int*)f->d_newMeas,
                   (int)newMeas.size(), (const int*)nullptr, f->d_newP);
        LAUNCH_CHECK(f, "new_landmark_kernel");
        if ((rc = remove_and_append(f, keep, addIds, addP, s.initialPointVariance, -1.0, true)) != EQVIO_OK) return rc;
        stage_mark(f, 2);
        return launch_correction(f, f->d_spec);
    }
    if (anyNew) {
        double depth = s.initialSceneDepth;
        if (s.useMedianDepth && P.h_gate) {  // getMedianSceneDepth, VIOFilter.cpp:366-380
            const double* depth2 = P.h_gate + 2 * N;
            std::vector<double> d2;
            for (int i = 0; i < N; ++i)
                if (keep[i]) d2.push_back(depth2[i]);
            if (!d2.empty()) {
                auto mid = d2.begin() + d2.size() / 2;
                std::nth_element(d2.begin(), mid, d2.end());
                depth = std::sqrt(*mid);
            }
        }
        for (int j = 0; j < n; ++j)
            if (!measInState[j]) {
                addIds.push_back(P.mids[j]);
                V3 b = cam_undistort(P.cam, P.my[2 * j], P.my[2 * j + 1]);
                addP.push_back(b.x * depth);
                addP.push_back(b.y * depth);
                addP.push_back(b.z * depth);
            }
    }
    if ((rc = remove_and_append(f, keep, addIds, addP, s.initialPointVariance, -1.0)) != EQVIO_OK) return rc;
    stage_mark(f, 2);
    return launch_correction(f, P.speculated ? f->d_spec : f->d_spec + 1);
}

// performVisionUpdate (VIO_eqf.cpp:105-135) on the measurement restricted to P.measKept.  Every kernel returns
// at once when *guard != 0.
int launch_correction(eqvio_filter* f, const int* guard) {
    auto& P = f->pend;
    const eqvio_settings& s = f->st;
    int rc;
    const int n = P.n;
    std::vector<int> kmids;
    std::vector<double> ky;
    for (int j = 0; j < n; ++j)
        if (P.measKept[j]) {
            kmids.push_back(P.mids[j]);
            ky.push_back(P.my[2 * j]);
            ky.push_back(P.my[2 * j + 1]);
        }
    const int nm = (int)kmids.size();
    P.corrected = false;
    if (nm == 0) {  // VIOFilter.cpp:223-224
        stage_mark(f, 3);
        return EQVIO_OK;
    }
    const int Nn = (int)f->ids.size();
    std::unordered_map<int, int> spos;
    spos.reserve(Nn * 2 + 1);
    for (int i = 0; i < Nn; ++i) spos[f->ids[i]] = i;
    std::vector<int> lmOf(nm), yIdx(nm);
    {
        std::vector<std::pair<int, int>> rows(nm);  // (state index, position in the kept measurement)
        for (int j = 0; j < nm; ++j) rows[j] = {spos.at(kmids[j]), j};
        if (f->corrMode != 1) std::sort(rows.begin(), rows.end());  // batch mode keeps the reference's ascending-id rows
        for (int j = 0; j < nm; ++j) {
            lmOf[j] = rows[j].first;
            yIdx[j] = rows[j].second;
        }
    }
    f->h_lmOfSorted = lmOf;
    if ((rc = upload(f, f->d_lmOf, lmOf.data(), nm)) != EQVIO_OK) return rc;
    if ((rc = upload(f, f->d_yIdx, yIdx.data(), nm)) != EQVIO_OK) return rc;
    if ((rc = upload(f, f->d_y, ky.data(), ky.size())) != EQVIO_OK) return rc;
    if ((rc = enqueue_correction(f, nm, guard)) != EQVIO_OK) return rc;
    stage_mark(f, 3);
    P.nStatus = 1 + Nn;
    if ((rc = download_async(f, &P.h_status, f->d_status, P.nStatus)) != EQVIO_OK) return rc;
    P.corrected = true;
    return EQVIO_OK;
}

// Launch plan of the lazy trailing updates (sequential chunks with look-ahead, EQVIO_TUNE_LAZY_DOWNDATE = M >= 1).  Pure host logic,
// exported as eqvio_plan_lazy_downdates so that the CPU tests can check its invariants without a GPU (tests/test_lazy_plan.py):
//   chunk c < nchunks-1:  factor(c);  [wait for the deferred launch of chunk waitRest];  urgent launch over the band [blo, bhi] of tile
//                         rows / columns chunk c+1 gathers from, up to chunk c;  if hasRest: deferred launch on the side stream over
//                         every lower tile with neither index in [xlo, xhi] (the bands of chunks c .. c+M), up to chunk c
//   chunk nchunks-1:      factor;  wait for the last deferred launch;  every lower tile up to chunk nchunks-1, mirrored.
struct LazyStep {
    int blo, bhi, nBand;  // urgent launch
    int waitRest;         // chunk whose deferred launch must have finished before this chunk's urgent / final launch (-1: none)
    int hasRest, xlo, xhi, nRest;
};
static void plan_lazy_downdates(int T, int nchunks, const int* blo, const int* bhi, int M, std::vector<LazyStep>& plan) {
    plan.assign(nchunks, LazyStep{0, 0, 0, -1, 0, 0, 0, 0});
    int lastRest = -1;
    for (int c = 0; c < nchunks; ++c) {
        LazyStep& s = plan[c];
        if (c == nchunks - 1) {
            s.blo = 0;
            s.bhi = T;
            s.nBand = T * (T + 1) / 2;
            s.waitRest = lastRest;
            break;
        }
        // the newest deferred launch this band's tiles were not left out of: largest r = kM - 1 <= c - M - 1
        if (c >= 2 * M) s.waitRest = ((c - M) / M) * M - 1;
        s.blo = blo[c];
        s.bhi = bhi[c];
        const int w = s.bhi - s.blo + 1;
        s.nBand = w * (T - 1 - s.bhi);
        for (int ti = s.blo; ti <= s.bhi; ++ti) s.nBand += ti + 1;
        if ((c + 1) % M == 0) {
            s.hasRest = 1;
            s.xlo = s.blo;
            s.xhi = bhi[std::min(c + M, nchunks - 2)];
            const int xw = s.xhi - s.xlo + 1;
            s.nRest = (T - xw) * (T - xw + 1) / 2;
            lastRest = c;
        }
    }
}

// One chunk factor launch (S_c, its elimination, Y_c, Gamma).
int launch_chunk_factor(eqvio_filter* f, int ldy, int dimp, int j0, int bc, double r2, const double* gin, double* gout, double* Yc,
                        const int* guard) {
    const int stage = f->stageS && f->haveSigMap ? 1 : 0;
    launch_pdl(f, chunk_factor_kernel, dim3(ldy / CH_COLS), dim3(CH_THREADS), (size_t)(stage ? CH_SMEM_STAGED : CH_SMEM_BASE), f->stream,
               (const double*)f->Sig[f->cur], f->ld, dimp, (const int*)f->d_lmOf, (const double*)f->d_Cblk, (const double*)f->d_ytilde, j0, bc, r2, gin, gout, Yc,
               f->d_status, guard, TL_SLOT(f), stage, f->sigMap[f->cur]);
    LAUNCH_CHECK(f, "chunk_factor_kernel");
    return EQVIO_OK;
}

// Block sweep with look-ahead (blockchol.cuh): stream A = f->stream carries the build and the chain of diagonal steps (programmatic
// edges between them), stream B = f->stream3 the panel / trailing kernels; diag(k) -> panel(k) -> trail(k) -> diag(k+2).
// While the per-kernel profile runs everything is issued on f->stream (event brackets around each class).
int enqueue_block_sweep(eqvio_filter* f, int nm, int dimp, double r2, const int* stateGuard, bool lateGate) {
    // lateGate: the gate flag is not known yet when the sweep starts (the gate runs beside it): kernels that only write scratch
    // (Z, panels, LT | XT) run unguarded, the trailing launches -- which write Sigma and Gamma -- wait for the gate and honour it
    const int* guard = lateGate ? f->d_spec + 1 : stateGuard;
    const int m = 2 * nm;
    const int nT = cdiv(m, BC_T);
    const int ldy = (dimp + BC_T - 1) / BC_T * BC_T, TW = ldy / BC_T;
    const int ldz = nT * BC_T + ldy;
    const bool serial = f->profiling;
    // A: build + diagonal chain; D: the two tiles diag(k+2) reads; B: the other tiles of the next block column; C: panel(k), the rest
    cudaStream_t sA = f->stream, sB = serial ? f->stream : f->stream4, sC = serial ? f->stream : f->stream3;
    while ((int)f->bcEv.size() < 4 * nT + 1) {
        cudaEvent_t e;
        CUDA_TRY(f, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        f->bcEv.push_back(e);
    }
    auto evDiag = [&](int k) { return f->bcEv[4 * k]; };
    auto evNext = [&](int k) { return f->bcEv[4 * k + 1]; };
    auto evRest = [&](int k) { return f->bcEv[4 * k + 2]; };
    auto evUrg = [&](int k) { return f->bcEv[4 * k + 3]; };
    const size_t zpElems = (size_t)(nT + TW) * YB_TILE;  // panels ping-pong: rest(k) still reads Zp(k) while panel(k+1) writes
    launch_pdl(f, bc_build_kernel, dim3(nT * (nT + 1) / 2 + TW * nT), dim3(256), (size_t)0, sA, (const double*)f->Sig[f->cur], f->ld, dimp,
               (const int*)f->d_lmOf, (const double*)f->d_Cblk, (const double*)f->d_ytilde, nm, r2, f->d_bcZ, ldz, nT, TW, f->d_bcCnt, f->d_bcMt, guard, TL_SLOT(f));
    LAUNCH_CHECK(f, "bc_build_kernel");
    for (int k = 0; k < nT; ++k) {
        // diag(k) needs two tiles of the trailing step k-2: counted off on the device (bc_diag_kernel), not an event edge into the chain
        const int waitCnt = (!serial && k >= 2) ? BC_TRAIL_URGENT : 0;
        int pk = prof_begin(f, PROF_BC_DIAG);
        launch_pdl(f, bc_diag_kernel, dim3(1), dim3(BC_DIAG_THREADS), (size_t)BC_DIAG_SMEM, sA, (const double*)f->d_bcZ, ldz, k, f->d_bcMt, f->d_status,
                   (const int*)f->d_bcCnt, waitCnt, guard, TL_SLOT(f));
        prof_end(f, pk);
        LAUNCH_CHECK(f, "bc_diag_kernel");
        if (!serial) CUDA_TRY(f, cudaEventRecord(evDiag(k), sA));
        double* Zp = f->d_bcZp + (size_t)(k & 1) * zpElems;
        const int below = nT + TW - k - 1;  // row tiles under the diagonal block (S, then W)
        const int q = nT - k - 1;
        if (serial) {
            int tk = prof_begin(f, PROF_BC_PANEL);
            bc_panel_kernel<<<2 * below, 128, BC_PANEL_SMEM, sB>>>(f->d_bcZ, ldz, k, f->d_bcMt, Zp, guard, TL_SLOT(f));
            prof_end(f, tk);
            LAUNCH_CHECK(f, "bc_panel_kernel");
            int sk = prof_begin(f, PROF_BC_TRAIL);
            bc_trail_kernel<<<2 * bc_trail_tiles(BC_PART_ALL, q, TW), DD_THREADS, DD_SMEM, sB>>>(f->d_bcZ, ldz, f->Sig[f->cur], f->ld, Zp, f->d_Gamma, stateGuard, k,
                                                                                               nT, TW, dimp, k == nT - 1 ? 1 : 0, (int)BC_PART_ALL,
                                                                                               f->d_bcCnt, TL_SLOT(f));
            prof_end(f, sk);
            LAUNCH_CHECK(f, "bc_trail_kernel");
        } else {
            const int nUrg = bc_trail_tiles(BC_PART_URGENT, q, TW), nNext = bc_trail_tiles(BC_PART_NEXT, q, TW);
            // chain stream, behind diag(k) as a programmatic dependent: T(k+2,k+1), T(k+2,k+2) -- their step k-1 came from next(k-1)
            if (nUrg > 0) {
                const int waitNext = k >= 1 ? 2 * bc_trail_tiles(BC_PART_NEXT, q + 1, TW) : 0;  // CTAs of next(k-1), counted off on the device
                launch_pdl(f, bc_next_kernel<4>, dim3(4 * nUrg), dim3(BC_NEXT_THREADS), (size_t)BC_NEXT_SMEM, sA, f->d_bcZ, ldz, (const double*)f->d_bcMt, guard, k,
                           nT, TW, (int)BC_PART_URGENT, f->d_bcCnt, waitNext, f->d_status, TL_SLOT(f));
                LAUNCH_CHECK(f, "bc_next_kernel<urgent>");
                CUDA_TRY(f, cudaEventRecord(evUrg(k), sA));
            }
            // B: the rest of block column k+1 and the urgent pair of step k+1 -- their step k-1 came from rest(k-1)
            if (nNext > 0) {
                CUDA_TRY(f, cudaStreamWaitEvent(sB, evDiag(k), 0));
                if (k >= 1) CUDA_TRY(f, cudaStreamWaitEvent(sB, evRest(k - 1), 0));
                bc_next_kernel<2><<<2 * nNext, BC_NEXT_THREADS, BC_NEXT_SMEM, sB>>>(f->d_bcZ, ldz, f->d_bcMt, guard, k, nT, TW, (int)BC_PART_NEXT, f->d_bcCnt,
                                                                                  0, f->d_status, TL_SLOT(f));
                LAUNCH_CHECK(f, "bc_next_kernel");
                CUDA_TRY(f, cudaEventRecord(evNext(k), sB));
            }
            // C: panels of every row tile (block column k is current through step k-1 after next(k-1) and urgent(k-1)), then all
            // the other trailing tiles
            CUDA_TRY(f, cudaStreamWaitEvent(sC, evDiag(k), 0));
            if (k == 0 && lateGate) CUDA_TRY(f, cudaStreamWaitEvent(sC, f->evGate, 0));
            if (k >= 1) {
                if (bc_trail_tiles(BC_PART_NEXT, q + 1, TW) > 0) CUDA_TRY(f, cudaStreamWaitEvent(sC, evNext(k - 1), 0));
                if (bc_trail_tiles(BC_PART_URGENT, q + 1, TW) > 0) CUDA_TRY(f, cudaStreamWaitEvent(sC, evUrg(k - 1), 0));
            }
            bc_panel_kernel<<<2 * below, 128, BC_PANEL_SMEM, sC>>>(f->d_bcZ, ldz, k, f->d_bcMt, Zp, guard, TL_SLOT(f));
            LAUNCH_CHECK(f, "bc_panel_kernel");
            bc_trail_kernel<<<2 * bc_trail_tiles(BC_PART_REST, q, TW), DD_THREADS, DD_SMEM, sC>>>(f->d_bcZ, ldz, f->Sig[f->cur], f->ld, Zp, f->d_Gamma, stateGuard, k,
                                                                                                nT, TW, dimp, k == nT - 1 ? 1 : 0, (int)BC_PART_REST,
                                                                                                f->d_bcCnt, TL_SLOT(f));
            LAUNCH_CHECK(f, "bc_trail_kernel<rest>");
            CUDA_TRY(f, cudaEventRecord(evRest(k), sC));
        }
    }
    if (!serial) {
        // everything joins the chain again: the last rest launch follows every next / urgent launch through the waits above, except
        // the ones of the last two steps
        CUDA_TRY(f, cudaStreamWaitEvent(sA, evRest(nT - 1), 0));
        for (int k = std::max(0, nT - 3); k < nT; ++k) {
            const int q = nT - k - 1;
            if (bc_trail_tiles(BC_PART_NEXT, q, TW) > 0) CUDA_TRY(f, cudaStreamWaitEvent(sA, evNext(k), 0));
        }
    }
    return EQVIO_OK;
}

// performVisionUpdate (VIO_eqf.cpp:105-135) in the symmetric form, for the nm measured landmarks whose pixels are
// in d_y and state indices in d_lmOf.  Every kernel returns at once when *guard != 0.
// fuseGate: the gate launch carries the measurement rows (gate_meas_kernel); fuseEst: the lift also emits the state estimate.
int enqueue_correction(eqvio_filter* f, int nm, const int* guard, bool fuseGate, bool fuseEst, bool lateGate) {
    auto& P = f->pend;
    (void)P;
    const eqvio_settings& s = f->st;
    const int Nn = (int)f->ids.size();
    const int m = 2 * nm;
    const int dimp = dimp_of(Nn);
    const int Mz = m + dimp + 1;
    const int ldz = (Mz + 7) & ~7;
    double* Z = f->d_Z;
    // the block sweep serves up to BC_MAX_ROWS measurement rows; larger updates (where the trailing work on W would dominate)
    // take the sequential chunks
    const bool blockSweep = f->corrMode == 2 && m <= BC_MAX_ROWS && !f->downdateTC;
    if (f->corrMode == 1) CUDA_TRY(f, cudaMemsetAsync(f->d_status, 0, (1 + (size_t)Nn) * sizeof(int), f->stream));
    const double r2 = s.measurementNoise * s.measurementNoise;
    const double* gammaFinal = f->d_Gamma;
    if (f->corrMode != 1) {
        // sequential chunks: see chunk_factor_kernel
        const int ldy = (dimp + 63) & ~63;
        const int T = ldy / DD_T;
        double* Y = f->d_Z;
        double* gin = f->d_Gamma;
        double* gout = f->d_Gamma2;
        // also clears the status words and Gamma (no memset nodes between the kernels of the update)
        if (lateGate) {
            // the rows were built behind the observer (enqueue_propagation); the gate runs on its own stream beside the sweep: only
            // the kernels that touch the filter state (trailing tiles of Sigma / Gamma, lift) look at its flag
            CUDA_TRY(f, cudaEventRecord(f->evG0, f->stream));
            CUDA_TRY(f, cudaStreamWaitEvent(f->stream5, f->evG0, 0));
            gate_fn(f)<<<cdiv(Nn, 128), 128, 0, f->stream5>>>(f->lm[f->lmcur], f->cap, Nn, f->Sig[f->cur], f->ld, f->d_measIdx, f->d_y, f->d_hdr,
                                                               s.coordinateChoice, f->d_gate, s.outlierThresholdAbs, s.outlierThresholdProb, f->d_spec,
                                                               TL_SLOT(f));
            LAUNCH_CHECK(f, "gate_kernel");
            CUDA_TRY(f, cudaEventRecord(f->evGate, f->stream5));
        } else if (fuseGate) {
            // one launch: gate CTAs (per state landmark) | measurement-row CTAs (per measured landmark); the rows are built
            // whatever the gate says -- d_spec + 1 is a constant 0
            const int gb = cdiv(Nn, 128), mb = cdiv(nm, 128);
            launch_pdl(f, gate_meas_kernel, dim3(gb + mb), dim3(128), (size_t)0, f->stream, gb, (const double*)f->lm[f->lmcur], f->cap, Nn,
                       (const double*)f->Sig[f->cur], f->ld, (const int*)f->d_measIdx, (const double*)f->d_y, (const FrameHeader*)f->d_hdr,
                       (int)s.coordinateChoice, f->d_gate, s.outlierThresholdAbs, s.outlierThresholdProb, f->d_spec, (const int*)f->d_lmOf, nm,
                       (const double*)f->d_y, s.useEquivariantOutput ? 1 : 0, f->d_Cblk, f->d_ytilde, 1, 0, (const int*)(f->d_spec + 1),
                       (const int*)f->d_yIdx, f->d_status, 1 + Nn, gin, dimp, (int*)nullptr, 0, TL_SLOT(f));
            LAUNCH_CHECK(f, "gate_meas_kernel");
        } else {
        meas_fn(f)<<<cdiv(nm, 128), 128, 0, f->stream>>>(f->lm[f->lmcur], f->cap, f->d_lmOf, nm, f->d_y, f->d_hdr, s.coordinateChoice,
                                                          s.useEquivariantOutput ? 1 : 0, f->d_Cblk, f->d_ytilde, 1, 0, guard, f->d_yIdx,
                                                          f->d_status, 1 + Nn, gin, dimp, (int*)nullptr, 0, TL_SLOT(f));
        LAUNCH_CHECK(f, "meas_kernel");
        }
        if (blockSweep) {
            const int rcb = enqueue_block_sweep(f, nm, dimp, r2, guard, lateGate);
            if (rcb != EQVIO_OK) return rcb;
            if (lateGate) CUDA_TRY(f, cudaStreamWaitEvent(f->stream, f->evGate, 0));
        } else {
        const int bcMax = std::max(1, std::min(f->chunkLm, CH_R / 2));
        const int nchunks = cdiv(nm, bcMax);
        const bool look = (f->lookahead == 1 || (f->lookahead == 2 && T >= 24)) && nchunks > 1 && !f->profiling && !f->downdateTC;
        if (look) {
            // Look-ahead: factor(c+1) only gathers the tile rows / columns of ITS landmarks (the band).  downdate(c) is split
            // into the band tiles (urgent, f->stream) and all other lower tiles (deferred, f->stream3, beside factor(c+1)).
            //   band(c)   after factor(c) [stream order] and rest(c-1) [event: both write the band of chunk c+1]
            //   rest(c)   after factor(c) [event] and rest(c-1) [stream order]
            //   factor(c+2) overwrites the Y buffer rest(c) reads: ordered through band(c+1)'s wait on rest(c)
            while ((int)f->chunkEv.size() < 2 * nchunks) {
                cudaEvent_t e;
                CUDA_TRY(f, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                f->chunkEv.push_back(e);
            }
            double* Ybuf[2] = {f->d_Z, f->d_Y2};
            const int* lo = f->h_lmOfSorted.data();
            const size_t chunkStride = (size_t)T * YB_TILE;
            const int M = f->lazyMerge;
            if (M >= 1 && nchunks > 2 && (size_t)nchunks * chunkStride <= f->zElems) {
                // Lazy trailing updates.  The panels of every chunk stay in d_Z; d_lvl[tile] = number of chunks applied to the tile.
                //   band(c)  [f->stream, after factor(c)]: the tiles chunk c+1 gathers from, brought up to chunk c (K = 64 .. 64 (M+1))
                //   rest(r)  [f->stream3, after factor(r)], r = M-1, 2M-1, ...: every other lower tile up to chunk r, EXCEPT the tile rows /
                //            columns of the bands r .. r+M, which the band launches beside it own; band(r+M+1) waits for rest(r)
                // so a deferred tile is read and written once per M chunks and the chain factor -> band -> factor never waits for a
                // rest launch that was issued less than M+1 chunks ago.
                std::vector<int> blo(nchunks, 0), bhi(nchunks, 0);
                for (int c = 0; c + 1 < nchunks; ++c) {
                    const int j1 = (c + 1) * bcMax;
                    const int nb = std::min(bcMax, nm - j1);
                    int rmin = lo[j1], rmax = rmin;
                    for (int q = 1; q < nb; ++q) {
                        rmin = std::min(rmin, lo[j1 + q]);
                        rmax = std::max(rmax, lo[j1 + q]);
                    }
                    blo[c] = (SOFF + 3 * rmin) / DD_T;
                    bhi[c] = (SOFF + 3 * rmax + 2) / DD_T;
                }
                CUDA_TRY(f, cudaMemsetAsync(f->d_lvl, 0, ((size_t)T * (T + 1) + nchunks) * sizeof(int), f->stream));
                std::vector<LazyStep> plan;
                plan_lazy_downdates(T, nchunks, blo.data(), bhi.data(), M, plan);
                for (int c = 0; c < nchunks; ++c) {
                    const LazyStep& st = plan[c];
                    const int j0 = c * bcMax;
                    const int bc = std::min(bcMax, nm - j0);
                    double* Yc = f->d_Z + (size_t)c * chunkStride;
                    cudaEvent_t evF = f->chunkEv[2 * c], evR = f->chunkEv[2 * c + 1];
                    f->pdlHold = (j0 == 0);
                    const int rc = launch_chunk_factor(f, ldy, dimp, j0, bc, r2, gin, gout, Yc, guard);
                    f->pdlHold = false;
                    if (rc != EQVIO_OK) return rc;
                    std::swap(gin, gout);
                    if (st.waitRest >= 0) CUDA_TRY(f, cudaStreamWaitEvent(f->stream, f->chunkEv[2 * st.waitRest + 1], 0));
                    if (c == nchunks - 1) {  // last chunk: every tile up to date, full symmetric storage again
                        chunk_downdate_kernel<false><<<st.nBand, DD_THREADS, DD_SMEM, f->stream>>>(
                            f->Sig[f->cur], f->Sig[f->cur], f->ld, f->d_Z, guard, 0, T, DD_ALL, T, TL_SLOT(f), f->d_lvl, nchunks, chunkStride);
                        LAUNCH_CHECK(f, "chunk_downdate_kernel");
                        break;
                    }
                    CUDA_TRY(f, cudaEventRecord(evF, f->stream));
                    if (f->bandSplit)
                        chunk_downdate_kernel<true><<<2 * st.nBand, DD_THREADS, DD_SMEM, f->stream>>>(
                            f->Sig[f->cur], f->Sig[f->cur], f->ld, f->d_Z, guard, st.blo, st.bhi, DD_BAND, T, TL_SLOT(f), f->d_lvl, c + 1, chunkStride);
                    else
                        chunk_downdate_kernel<false><<<st.nBand, DD_THREADS, DD_SMEM, f->stream>>>(
                            f->Sig[f->cur], f->Sig[f->cur], f->ld, f->d_Z, guard, st.blo, st.bhi, DD_BAND, T, TL_SLOT(f), f->d_lvl, c + 1, chunkStride);
                    LAUNCH_CHECK(f, "chunk_downdate_kernel");
                    if (st.hasRest) {
                        if (f->restAfterBand) CUDA_TRY(f, cudaEventRecord(evF, f->stream));
                        CUDA_TRY(f, cudaStreamWaitEvent(f->stream3, evF, 0));
                        if (st.nRest > 0) {
                            f->pdlHold = true;
                            const int perSm = f->restPersist;  // 0: one CTA per tile
                            const int grid = perSm ? std::min(st.nRest, perSm * f->smCount) : st.nRest;
                            launch_pdl(f, chunk_downdate_kernel<false>, dim3(grid), dim3(DD_THREADS), (size_t)(perSm ? DD_SMEM_PERSIST : DD_SMEM),
                                       f->stream3, (const double*)f->Sig[f->cur], f->Sig[f->cur], f->ld, (const double*)f->d_Z, guard, st.xlo, st.xhi,
                                       (int)DD_REST, T, TL_SLOT(f), f->d_lvl, c + 1, chunkStride,
                                       perSm ? f->d_lvl + (size_t)T * (T + 1) + c : (int*)nullptr, st.nRest);
                            f->pdlHold = false;
                            LAUNCH_CHECK(f, "chunk_downdate_kernel");
                        }
                        CUDA_TRY(f, cudaEventRecord(evR, f->stream3));
                    }
                }
            } else
            for (int c = 0; c < nchunks; ++c) {
                const int j0 = c * bcMax;
                const int bc = std::min(bcMax, nm - j0);
                double* Yc = Ybuf[c & 1];
                cudaEvent_t evF = f->chunkEv[2 * c], evR = f->chunkEv[2 * c + 1];
                f->pdlHold = (j0 == 0);  // chunk 0 follows meas_kernel, whose output the kernel stages ahead of its dependency wait
                const int rc = launch_chunk_factor(f, ldy, dimp, j0, bc, r2, gin, gout, Yc, guard);
                f->pdlHold = false;
                if (rc != EQVIO_OK) return rc;
                std::swap(gin, gout);
                if (c > 0) CUDA_TRY(f, cudaStreamWaitEvent(f->stream, f->chunkEv[2 * (c - 1) + 1], 0));  // rest(c-1) done
                if (c == nchunks - 1) {  // last chunk: everything, and full symmetric storage again
                    chunk_downdate_kernel<false><<<T * (T + 1) / 2, DD_THREADS, DD_SMEM, f->stream>>>(f->Sig[f->cur], f->Sig[f->cur], f->ld, Yc, guard,
                                                                                              0, T, DD_ALL, T, TL_SLOT(f));
                    LAUNCH_CHECK(f, "chunk_downdate_kernel");
                    break;
                }
                const int nb = std::min(bcMax, nm - (j0 + bcMax));
                int rmin = lo[j0 + bcMax], rmax = rmin;
                for (int q = 1; q < nb; ++q) {
                    rmin = std::min(rmin, lo[j0 + bcMax + q]);
                    rmax = std::max(rmax, lo[j0 + bcMax + q]);
                }
                const int mlo = (SOFF + 3 * rmin) / DD_T, mhi = (SOFF + 3 * rmax + 2) / DD_T;
                const int w = mhi - mlo + 1;
                int nBand = w * (T - 1 - mhi);
                for (int ti = mlo; ti <= mhi; ++ti) nBand += ti + 1;
                const int nRest = (T - w) * (T - w + 1) / 2;
                CUDA_TRY(f, cudaEventRecord(evF, f->stream));
                chunk_downdate_kernel<false><<<nBand, DD_THREADS, DD_SMEM, f->stream>>>(f->Sig[f->cur], f->Sig[f->cur], f->ld, Yc, guard, mlo, mhi,
                                                                                 DD_BAND, T, TL_SLOT(f));
                LAUNCH_CHECK(f, "chunk_downdate_kernel");
                if (f->restAfterBand) CUDA_TRY(f, cudaEventRecord(evF, f->stream));
                CUDA_TRY(f, cudaStreamWaitEvent(f->stream3, evF, 0));
                if (nRest > 0) {
                    f->pdlHold = true;
                    launch_pdl(f, chunk_downdate_kernel<false>, dim3(nRest), dim3(DD_THREADS), (size_t)DD_SMEM, f->stream3, (const double*)f->Sig[f->cur],
                               f->Sig[f->cur], f->ld, (const double*)Yc, guard, mlo, mhi, (int)DD_REST, T, TL_SLOT(f), (int*)nullptr, 1, (size_t)0, (int*)nullptr, 0);
                    f->pdlHold = false;
                    LAUNCH_CHECK(f, "chunk_downdate_kernel");
                }
                CUDA_TRY(f, cudaEventRecord(evR, f->stream3));
            }
        } else {
            for (int j0 = 0; j0 < nm; j0 += bcMax) {
                const int bc = std::min(bcMax, nm - j0);
                int pk = prof_begin(f, PROF_PANEL);
                f->pdlHold = (j0 == 0);  // chunk 0 follows meas_kernel, whose output the kernel stages ahead of its dependency wait
                const int rc = launch_chunk_factor(f, ldy, dimp, j0, bc, r2, gin, gout, Y, guard);
                prof_end(f, pk);
                f->pdlHold = false;
                if (rc != EQVIO_OK) return rc;
                int sk = prof_begin(f, PROF_SYRK);
                if (f->downdateTC) {
                    const int ncols = cdiv(ldy, TC_T) * TC_T;  // Y's pad columns up to ld are zero
                    const int T128 = ncols / TC_T;
                    if (ncols > ldy)  // the last 128-column block is only half covered by Y: the rest must read as zeros
                        CUDA_TRY(f, cudaMemsetAsync(f->d_Ysplit + (size_t)(T128 - 1) * TC_BLOCK_BYTES, 0, TC_BLOCK_BYTES, f->stream));
                    y_split_kernel<<<cdiv(ldy * 8, 256), 256, 0, f->stream>>>(Y, f->d_Ysplit, ldy);
                    LAUNCH_CHECK(f, "y_split_kernel");
                    chunk_downdate_tc_kernel<<<T128 * (T128 + 1) / 2, 128, TC_SMEM, f->stream>>>(f->Sig[f->cur], f->Sig[f->cur], f->ld,
                                                                                                 f->d_Ysplit, guard);
                    LAUNCH_CHECK(f, "chunk_downdate_tc_kernel");
                } else {
                    // mirror tiles are only refreshed where the next chunk will gather (rows of its landmarks); the last
                    // chunk writes them all
                    int mlo = 0, mhi = T;
                    if (j0 + bcMax < nm && f->lazyMirror) {
                        const int* lo = f->h_lmOfSorted.data();
                        const int nb = std::min(bcMax, nm - (j0 + bcMax));
                        int rmin = lo[j0 + bcMax], rmax = lo[j0 + bcMax];
                        for (int q = 1; q < nb; ++q) {
                            rmin = std::min(rmin, lo[j0 + bcMax + q]);
                            rmax = std::max(rmax, lo[j0 + bcMax + q]);
                        }
                        mlo = (SOFF + 3 * rmin) / DD_T;
                        mhi = (SOFF + 3 * rmax + 2) / DD_T;
                    }
                    // one wave with SMs to spare: two CTAs per tile (the launch lasts as long as its slowest CTA)
                    if (f->splitDowndate && T * (T + 1) / 2 <= f->smCount)
                        launch_pdl(f, chunk_downdate_kernel<true>, dim3(T * (T + 1)), dim3(DD_THREADS), DD_SMEM, f->stream, f->Sig[f->cur],
                                   f->Sig[f->cur], f->ld, Y, guard, mlo, mhi, (int)DD_ALL, T, TL_SLOT(f), (int*)nullptr, 1, (size_t)0, (int*)nullptr, 0);
                    else
                        launch_pdl(f, chunk_downdate_kernel<false>, dim3(T * (T + 1) / 2), dim3(DD_THREADS), DD_SMEM, f->stream, f->Sig[f->cur],
                                   f->Sig[f->cur], f->ld, Y, guard, mlo, mhi, (int)DD_ALL, T, TL_SLOT(f), (int*)nullptr, 1, (size_t)0, (int*)nullptr, 0);
                    LAUNCH_CHECK(f, "chunk_downdate_kernel");
                }
                prof_end(f, sk);
                std::swap(gin, gout);
            }
        }
        gammaFinal = gin;
        }
    } else {
    meas_fn(f)<<<cdiv(nm, 128), 128, 0, f->stream>>>(f->lm[f->lmcur], f->cap, f->d_lmOf, nm, f->d_y, f->d_hdr, s.coordinateChoice,
                                                      s.useEquivariantOutput ? 1 : 0, f->d_Cblk, Z, ldz, m + dimp, guard, f->d_yIdx,
                                                      nullptr, 0, nullptr, 0, nullptr, 0, TL_SLOT(f));
    LAUNCH_CHECK(f, "meas_kernel");
    zbuild_kernel<<<dim3(cdiv(dimp, 256), nm), 256, 0, f->stream>>>(f->Sig[f->cur], f->ld, dimp, f->d_lmOf, f->d_Cblk, Z, ldz, m);
    LAUNCH_CHECK(f, "zbuild_kernel");
    sbuild_kernel<<<dim3(cdiv(m, 128), m), 128, 0, f->stream>>>(f->d_lmOf, f->d_Cblk, Z, ldz, m, r2);
    LAUNCH_CHECK(f, "sbuild_kernel");
    for (int k = 0; k < m; k += NB) {
        const int nbk = std::min(NB, m - k);
        const int below = Mz - (k + nbk);
        int pk = prof_begin(f, PROF_PANEL);
        chol_panel_kernel<<<std::max(1, cdiv(below, PANEL_THREADS)), PANEL_THREADS, 0, f->stream>>>(Z, ldz, Mz, k, nbk,
                                                                                                     f->d_status, f->d_Lout);
        prof_end(f, pk);
        LAUNCH_CHECK(f, "chol_panel_kernel");
        if (k + nbk < m) {
            const int o = k + nbk;
            const int Mr = Mz - o, Nc = m - o;
            int tk = prof_begin(f, PROF_TRAIL);
            gemm_nt_sub_kernel<false><<<dim3(cdiv(Nc, GBN), cdiv(Mr, GBM)), 128, 0, f->stream>>>(
                Z + (size_t)o * ldz + o, ldz, Z + (size_t)k * ldz + o, ldz, Z + (size_t)k * ldz + o, ldz, Mr, Nc, nbk);
            prof_end(f, tk);
            LAUNCH_CHECK(f, "gemm_nt_sub_kernel<trail>");
        }
    }
    {
        int sk = prof_begin(f, PROF_SYRK);
        gemm_nt_sub_kernel<true><<<dim3(cdiv(dimp, GBN), cdiv(dimp, GBM)), 128, 0, f->stream>>>(f->Sig[f->cur], f->ld, Z + m, ldz,
                                                                                                Z + m, ldz, dimp, dimp, m);
        prof_end(f, sk);
        LAUNCH_CHECK(f, "gemm_nt_sub_kernel<syrk>");
    }
    gamma_kernel<<<cdiv(dimp, 128), 128, m * sizeof(double), f->stream>>>(Z, ldz, m, dimp, f->d_Gamma);
    LAUNCH_CHECK(f, "gamma_kernel");
    }
    launch_pdl(f, lift_fn(f), dim3(cdiv(std::max(Nn, 1), 128)), dim3(128), (size_t)(0), f->stream, f->lm[f->lmcur], f->cap, Nn, f->d_xi0s, f->d_Xs[f->xcur], gammaFinal,
                                                                   s.useDiscreteInnovationLift ? 1 : 0, s.coordinateChoice,
                                                                   f->d_status, f->d_status + 1, guard, fuseEst ? f->d_out : (double*)nullptr,
                                                                   s.coordinateChoice == EQVIO_COORD_NORMAL ? (const double*)(f->d_normalM + 441) : (const double*)nullptr, TL_SLOT(f));
    LAUNCH_CHECK(f, "lift_kernel");
    return EQVIO_OK;
}

// ---- phase C: wait, check the device status word, drop invalid landmarks ---------------------------
int vision_phase_c(eqvio_filter* f, int* did_update) {
    auto& P = f->pend;
    if (did_update) *did_update = 0;
    {
        const auto w0 = std::chrono::steady_clock::now();
        CUDA_TRY(f, cudaStreamSynchronize(f->stream));
        f->hostWaitUs = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - w0).count();
    }
    prof_collect(f);
    if (f->stageTiming && P.active) {
        for (int i = 0; i < 3; ++i) f->stageMs[i] = 0;
        if (P.steady && !f->steadySplit) {  // a replayed graph: one bracket around the whole update
            float ms = 0;
            if (cudaEventElapsedTime(&ms, f->stageEv[0], f->stageEv[3]) == cudaSuccess) f->stageMs[2] = ms;
        } else {
            for (int i = 0; i < 3; ++i) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, f->stageEv[i], f->stageEv[i + 1]) == cudaSuccess) f->stageMs[i] = ms;
            }
        }
        if (f->augTimed) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, f->augEv[0], f->augEv[1]) == cudaSuccess) f->stageMs[1] += ms;
            f->augTimed = false;
        }
    }
    const bool wasSteady = P.active && P.steady;
    bool redone = false;
    if (P.active && P.speculated && !P.ignoreGate && P.h_spec && *P.h_spec != 0) {
        redone = true;
        // a gate tripped: the guarded correction did nothing.  Decide exactly (the gate scalars are on the host
        // by now), remove the outliers from the already lost-compacted state and correct without a guard.
        P.speculated = false;
        std::vector<char> outlier;
        decide_outliers(f, true, outlier);
        std::vector<char> keepNow;
        for (size_t i = 0; i < P.oldIds.size(); ++i)
            if (P.keep[i]) keepNow.push_back(!outlier[i]);
        std::vector<int> addIds;
        std::vector<double> addP;
        if (P.specNewCount > 0) {
            // the speculative append used the median depth of ALL kept landmarks: drop those landmarks and add them again
            // with the depth of the landmarks that survive the gate (VIOFilter.cpp:206-219 order: outliers first, then new ids)
            for (int k = 0; k < P.specNewCount; ++k) keepNow.push_back(0);
            const eqvio_settings& s = f->st;
            const int N0 = (int)P.oldIds.size();
            double depth = s.initialSceneDepth;
            if (s.useMedianDepth && P.h_gate) {
                const double* depth2 = P.h_gate + 2 * N0;
                std::vector<double> d2;
                for (int i = 0; i < N0; ++i)
                    if (P.keep[i] && !outlier[i]) d2.push_back(depth2[i]);
                if (!d2.empty()) {
                    auto mid = d2.begin() + d2.size() / 2;
                    std::nth_element(d2.begin(), mid, d2.end());
                    depth = std::sqrt(*mid);
                }
            }
            std::vector<char> inState(P.n, 0);
            for (int i = 0; i < N0; ++i)
                if (P.measIdx[i] >= 0) inState[P.measIdx[i]] = 1;
            for (int j = 0; j < P.n; ++j)
                if (!inState[j]) {
                    addIds.push_back(P.mids[j]);
                    V3 b = cam_undistort(P.cam, P.my[2 * j], P.my[2 * j + 1]);
                    addP.push_back(b.x * depth);
                    addP.push_back(b.y * depth);
                    addP.push_back(b.z * depth);
                }
        }
        int rc = remove_and_append(f, keepNow, addIds, addP, f->st.initialPointVariance, -1.0);
        if (rc == EQVIO_OK) rc = launch_correction(f, f->d_spec + 1);
        if (rc != EQVIO_OK) return rc;
        CUDA_TRY(f, cudaStreamSynchronize(f->stream));
        prof_collect(f);
    }
    if (!P.active || !P.corrected) {
        P.active = false;
        return EQVIO_OK;
    }
    P.active = false;
    const int st = P.h_status[0];
    if (st & 1) {
        f->err = "innovation covariance S is not positive definite";
        return EQVIO_ERR_NUMERIC;
    }
    if (st & 2) {
        f->err = "NaN detected in the correction";
        return EQVIO_ERR_NUMERIC;
    }
    if (st & 8) {
        f->err = "chunk factor kernel: a bounded device-side wait ran out";
        return EQVIO_ERR_CUDA;
    }
    if (did_update) *did_update = 1;
    f->estValid = wasSteady && !redone && !(st & 4);
    if (st & 4) {  // removeInvalidLandmarks, VIO_eqf.cpp:213-223
        const int Nn = (int)f->ids.size();
        std::vector<char> keep(Nn, 1);
        for (int i = 0; i < Nn; ++i)
            if (P.h_status[1 + i]) keep[i] = 0;
        int rc = remove_and_append(f, keep, {}, {}, 0.0, -1.0);
        if (rc != EQVIO_OK) return rc;
        CUDA_TRY(f, cudaStreamSynchronize(f->stream));
    }
    return EQVIO_OK;
}

int make_filter(const eqvio_settings* s, int device, int capacity, void* stream, eqvio_filter** out) {
    if (!s || !out || capacity < 0) {
        g_createError = "invalid arguments";
        return EQVIO_ERR_INVALID_ARG;
    }
    *out = nullptr;
    if (s->coordinateChoice != EQVIO_COORD_EUCLIDEAN && s->coordinateChoice != EQVIO_COORD_INVDEPTH &&
        s->coordinateChoice != EQVIO_COORD_NORMAL) {
        g_createError = "coordinateChoice must be Euclidean, InvDepth or Normal";
        return EQVIO_ERR_UNSUPPORTED;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_createError = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return EQVIO_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        g_createError = "device index out of range";
        return EQVIO_ERR_INVALID_ARG;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_createError = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return EQVIO_ERR_CUDA;
    }
    e = cudaFuncSetAttribute(chunk_downdate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_SMEM_PERSIST);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chunk_downdate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chunk_downdate_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // factor CTAs of the next chunk must fit beside the deferred downdate CTAs: keep the shared-memory carve-out at its maximum
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chunk_downdate_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chunk_factor_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chunk_downdate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bc_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_DIAG_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bc_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_PANEL_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bc_trail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bc_next_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_NEXT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bc_next_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_NEXT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chunk_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(ChunkSmem) > (size_t)CH_SMEM_STAGED ? sizeof(ChunkSmem) : (size_t)CH_SMEM_STAGED));
    if (e != cudaSuccess) {
        g_createError = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
        return EQVIO_ERR_CUDA;
    }
    eqvio_filter* f = new eqvio_filter();
    f->st = *s;
    f->device = device;
    f->cap = capacity;
    if (const char* pf = std::getenv("EQVIO_B200_PREFETCH")) f->prefetchSigma = std::atoi(pf) != 0;
    if (const char* er = std::getenv("EQVIO_B200_EARLY_ROWS")) f->earlyRows = std::atoi(er) != 0;
    if (const char* lz = std::getenv("EQVIO_B200_LAZY")) f->lazyMerge = std::max(0, std::min(8, std::atoi(lz)));
    if (const char* bs = std::getenv("EQVIO_B200_BAND_SPLIT")) f->bandSplit = std::atoi(bs) != 0;
    if (const char* rb = std::getenv("EQVIO_B200_REST_AFTER_BAND")) f->restAfterBand = std::atoi(rb) != 0;
    if (const char* rp = std::getenv("EQVIO_B200_REST_PERSIST")) f->restPersist = std::max(0, std::min(4, std::atoi(rp)));
    if (const char* cm = std::getenv("EQVIO_B200_CORRECTION")) {  // A/B runs of whole test / bench commands
        const int v = std::atoi(cm);
        if (v >= 0 && v <= 2) f->corrMode = v;
    }
    if (cudaDeviceGetAttribute(&f->smCount, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || f->smCount <= 0) f->smCount = 148;
    if (stream) {
        f->stream = static_cast<cudaStream_t>(stream);
    } else {
        int prLeast = 0, prGreatest = 0;
        e = cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&f->stream, cudaStreamNonBlocking, prGreatest);
        if (e != cudaSuccess) {
            g_createError = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
            delete f;
            return EQVIO_ERR_CUDA;
        }
        f->ownStream = true;
    }
    int rc = alloc_device(f);
    if (rc != EQVIO_OK) {
        g_createError = f->err;
        eqvio_destroy(f);
        return rc;
    }
    *out = f;
    return EQVIO_OK;
}

#define ENTER(f)                                        \
    if (!(f)) return EQVIO_ERR_INVALID_ARG;             \
    {                                                   \
        cudaError_t e_ = cudaSetDevice((f)->device);    \
        if (e_ != cudaSuccess) {                        \
            (f)->err = cudaGetErrorString(e_);          \
            return EQVIO_ERR_CUDA;                      \
        }                                               \
    }

}  // namespace

extern "C" {

void eqvio_settings_default(eqvio_settings* s) {  // VIOFilterSettings.h:59-98
    if (!s) return;
    s->biasOmegaProcessVariance = s->biasAccelProcessVariance = s->attitudeProcessVariance = s->positionProcessVariance =
        s->velocityProcessVariance = s->cameraAttitudeProcessVariance = s->cameraPositionProcessVariance =
            s->pointProcessVariance = 0.001;
    s->velGyrNoise = 1e-4;
    s->velAccNoise = 1e-3;
    s->velGyrBiasWalk = 1e-5;
    s->velAccBiasWalk = 1e-3;
    s->measurementNoise = 2.0;
    s->outlierThresholdAbs = 1e8;
    s->outlierThresholdProb = 1e8;
    s->featureRetention = 0.3;
    s->initialAttitudeVariance = 1.0e-4;
    s->initialPositionVariance = 1.0e-4;
    s->initialVelocityVariance = 1.0e-2;
    s->initialCameraAttitudeVariance = 1.0e-5;
    s->initialCameraPositionVariance = 1.0e-4;
    s->initialPointVariance = 1.0;
    s->initialPointDepthVariance = -1.0;
    s->initialBiasOmegaVariance = 0.1;
    s->initialBiasAccelVariance = 0.1;
    s->initialSceneDepth = 1.0;
    s->useDiscreteInnovationLift = 1;
    s->useDiscreteVelocityLift = 1;
    s->useDiscreteStateMatrix = 0;
    s->fastRiccati = 0;
    s->useMedianDepth = 1;
    s->useFeaturePredictions = 0;
    s->useEquivariantOutput = 1;
    s->removeLostLandmarks = 1;
    s->coordinateChoice = EQVIO_COORD_EUCLIDEAN;
    s->cameraOffset[0] = 1.0;
    for (int i = 1; i < 7; ++i) s->cameraOffset[i] = 0.0;
}

// StandardCamera::computeInverseDistortion (StandardCamera.cpp:113-145): least-squares fit of five
// inverse coefficients on a grid of normalised points, solved by Householder QR.
int eqvio_camera_fit_inverse_distortion(eqvio_camera* cam) {
    if (!cam) return EQVIO_ERR_INVALID_ARG;
    int w = cam->width, h = cam->height;
    if ((long long)w * h == 0) {
        w = (int)std::lround(cam->cx * 2);
        h = (int)std::lround(cam->cy * 2);
    }
    const int maxPoints = 30;
    const int sx = w / maxPoints, sy = h / maxPoints;
    if (sx <= 0 || sy <= 0) return EQVIO_ERR_INVALID_ARG;
    std::vector<double> A, b;  // row-major rows x 5
    for (int x = 0; x < w; x += sx)
        for (int y = 0; y < h; y += sy) {
            const double nx = (x - cam->cx) / cam->fx, ny = (y - cam->cy) / cam->fy;
            double px, py;
            distort_homogeneous(nx, ny, cam->dist, cam->ndist, px, py);
            const double r2 = px * px + py * py;
            const double r0[5] = {px * r2, px * r2 * r2, 2 * px * py, r2 + 2 * px * px, px * r2 * r2 * r2};
            const double r1[5] = {py * r2, py * r2 * r2, r2 + 2 * py * py, 2 * px * py, py * r2 * r2 * r2};
            A.insert(A.end(), r0, r0 + 5);
            A.insert(A.end(), r1, r1 + 5);
            b.push_back(nx - px);
            b.push_back(ny - py);
        }
    const int rows = (int)b.size();
    if (rows < 5) return EQVIO_ERR_INVALID_ARG;
    // Householder QR with column pivoting on [A | b]
    int perm[5] = {0, 1, 2, 3, 4};
    for (int k = 0; k < 5; ++k) {
        int best = k;
        double bestN = -1;
        for (int c = k; c < 5; ++c) {
            double s = 0;
            for (int r = k; r < rows; ++r) s += A[r * 5 + c] * A[r * 5 + c];
            if (s > bestN) {
                bestN = s;
                best = c;
            }
        }
        if (best != k) {
            for (int r = 0; r < rows; ++r) std::swap(A[r * 5 + k], A[r * 5 + best]);
            std::swap(perm[k], perm[best]);
        }
        double nrm = std::sqrt(bestN);
        if (nrm == 0) continue;
        const double alpha = A[k * 5 + k] > 0 ? -nrm : nrm;
        std::vector<double> v(rows - k);
        for (int r = k; r < rows; ++r) v[r - k] = A[r * 5 + k];
        v[0] -= alpha;
        double vn = 0;
        for (double t : v) vn += t * t;
        if (vn == 0) continue;
        for (int c = k; c < 5; ++c) {
            double d = 0;
            for (int r = k; r < rows; ++r) d += v[r - k] * A[r * 5 + c];
            d = 2 * d / vn;
            for (int r = k; r < rows; ++r) A[r * 5 + c] -= d * v[r - k];
        }
        double d = 0;
        for (int r = k; r < rows; ++r) d += v[r - k] * b[r];
        d = 2 * d / vn;
        for (int r = k; r < rows; ++r) b[r] -= d * v[r - k];
    }
    double xs[5] = {0, 0, 0, 0, 0};
    for (int k = 4; k >= 0; --k) {
        double s = b[k];
        for (int c = k + 1; c < 5; ++c) s -= A[k * 5 + c] * xs[c];
        xs[k] = (A[k * 5 + k] != 0) ? s / A[k * 5 + k] : 0.0;
    }
    for (int k = 0; k < 5; ++k) cam->inv_dist[perm[k]] = xs[k];
    return EQVIO_OK;
}

int eqvio_create(const eqvio_settings* s, int device, int capacity, void* stream_or_null, eqvio_filter** out) {
    int rc = make_filter(s, device, capacity, stream_or_null, out);
    if (rc != EQVIO_OK) return rc;
    eqvio_filter* f = *out;
    double sensor[23];
    for (int i = 0; i < 23; ++i) sensor[i] = 0.0;
    sensor[6] = 1.0;  // identity pose
    for (int i = 0; i < 7; ++i) sensor[16 + i] = s->cameraOffset[i];
    std::vector<double> diag;
    initial_diag(f->st, 0, true, diag);
    rc = reset_state(f, sensor, 0, nullptr, nullptr, diag);
    if (rc != EQVIO_OK) {
        g_createError = f->err;
        eqvio_destroy(f);
        *out = nullptr;
        return rc;
    }
    f->time = -1.0;
    f->initialised = false;
    return EQVIO_OK;
}

int eqvio_create_from_state(const eqvio_settings* s, int device, int capacity, void* stream_or_null, const double sensor[23],
                            int n, const int* ids, const double* p, double time, eqvio_filter** out) {
    if (!sensor || n < 0 || (n > 0 && (!ids || !p))) {
        g_createError = "invalid arguments";
        return EQVIO_ERR_INVALID_ARG;
    }
    if (n > capacity) {
        g_createError = "initial landmark count exceeds capacity";
        return EQVIO_ERR_CAPACITY;
    }
    int rc = make_filter(s, device, capacity, stream_or_null, out);
    if (rc != EQVIO_OK) return rc;
    eqvio_filter* f = *out;
    std::vector<double> diag;
    initial_diag(f->st, n, true, diag);
    rc = reset_state(f, sensor, n, ids, p, diag);
    if (rc != EQVIO_OK) {
        g_createError = f->err;
        eqvio_destroy(f);
        *out = nullptr;
        return rc;
    }
    f->time = time;
    f->initialised = true;
    return EQVIO_OK;
}

void eqvio_destroy(eqvio_filter* f) {
    if (!f) return;
    cudaSetDevice(f->device);
    if (f->stream) cudaStreamSynchronize(f->stream);
    for (int k = 0; k < 2; ++k) {
        cudaFree(f->Sig[k]);
        cudaFree(f->lm[k]);
        cudaFree(f->dids[k]);
    }
    cudaFree(f->d_xi0s);
    cudaFree(f->d_Xs[0]);
    cudaFree(f->d_Xs[1]);
    if (f->stream2) cudaStreamDestroy(f->stream2);
    if (f->stream3) cudaStreamDestroy(f->stream3);
    if (f->stream4) cudaStreamDestroy(f->stream4);
    if (f->stream5) cudaStreamDestroy(f->stream5);
    if (f->evFork) cudaEventDestroy(f->evFork);
    if (f->evJoin) cudaEventDestroy(f->evJoin);
    if (f->evG0) cudaEventDestroy(f->evG0);
    if (f->evGate) cudaEventDestroy(f->evGate);
    if (f->evStrip0) cudaEventDestroy(f->evStrip0);
    if (f->evStrip1) cudaEventDestroy(f->evStrip1);
    cudaFree(f->d_ctx);
    cudaFree(f->d_steps);
    cudaFree(f->d_rows);
    cudaFree(f->d_uv);
    cudaFree(f->d_Z);
    cudaFree(f->d_Lout);
    cudaFree(f->d_bcZ);
    cudaFree(f->d_bcZp);
    cudaFree(f->d_bcMt);
    cudaFree(f->d_bcCnt);
    for (auto& e : f->bcEv) cudaEventDestroy(e);
    cudaFree(f->d_Y2);
    cudaFree(f->d_lvl);
    cudaFree(f->d_Ysplit);
    for (auto& e : f->chunkEv) cudaEventDestroy(e);
    cudaFree(f->d_Cblk);
    cudaFree(f->d_Gamma);
    cudaFree(f->d_Gamma2);
    cudaFree(f->d_ytilde);
    cudaFree(f->d_frame);
    cudaFree(f->d_hdrSteps);
    f->sric.release();
    if (f->h_frame) cudaFreeHost(f->h_frame);
    if (f->h_out) cudaFreeHost(f->h_out);
    cudaFree(f->d_outblk);
    cudaFree(f->d_normalM);
    cudaFree(f->d_keepI);
    cudaFree(f->d_newMeas);
    for (auto& g : f->graphs)
        if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    cudaFree(f->d_mapblk);
    for (int i = 0; i < 2; ++i)
        if (f->augEv[i]) cudaEventDestroy(f->augEv[i]);
    for (auto& A : f->arenaSets) {
        for (auto& a : A.blocks) cudaFreeHost(a.first);
        if (A.ev) cudaEventDestroy(A.ev);
    }
    for (auto& p : f->evPool) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    for (int i = 0; i < 4; ++i)
        if (f->stageEv[i]) cudaEventDestroy(f->stageEv[i]);
    if (f->ownStream && f->stream) cudaStreamDestroy(f->stream);
    delete f;
}

const char* eqvio_last_error(const eqvio_filter* f) { return f ? f->err.c_str() : g_createError.c_str(); }

int eqvio_initialise_from_imu(eqvio_filter* f, double stamp, const double gyr[3], const double acc[3]) {
    ENTER(f);
    f->estValid = false;
    (void)gyr;
    if (!acc) return EQVIO_ERR_INVALID_ARG;
    stage_reset(f);
    for (int i = 0; i < 6; ++i) f->xi0s[i] = 0.0;
    Quat q = quat_from_two_vectors(normalized(V3{acc[0], acc[1], acc[2]}), V3{0, 0, 1});
    f->xi0s[6] = q.w;
    f->xi0s[7] = q.x;
    f->xi0s[8] = q.y;
    f->xi0s[9] = q.z;
    for (int i = 10; i < 16; ++i) f->xi0s[i] = 0.0;
    int rc = upload(f, f->d_xi0s, f->xi0s, 23);
    f->normalMValid = false;
    if (rc != EQVIO_OK) return rc;
    CUDA_TRY(f, cudaStreamSynchronize(f->stream));
    f->initialised = true;
    f->time = stamp;
    return EQVIO_OK;
}

int eqvio_set_state(eqvio_filter* f, const double sensor[23], int n, const int* ids, const double* p) {
    ENTER(f);
    f->estValid = false;
    if (!sensor || n < 0 || (n > 0 && (!ids || !p))) return EQVIO_ERR_INVALID_ARG;
    stage_reset(f);
    std::vector<double> diag;
    initial_diag(f->st, n, false, diag);
    int rc = reset_state(f, sensor, n, ids, p, diag);
    if (rc != EQVIO_OK) return rc;
    f->initialised = true;
    return EQVIO_OK;
}

int eqvio_set_landmarks(eqvio_filter* f, int n, const int* ids, const double* p) {
    ENTER(f);
    f->estValid = false;
    if (n < 0 || (n > 0 && (!ids || !p))) return EQVIO_ERR_INVALID_ARG;
    if (n != (int)f->ids.size()) {
        // the reference overwrites a block of an unresized Sigma (VIOFilter.cpp:94-101); any other
        // count leaves its state inconsistent, so it is rejected here
        f->err = "set_landmarks: count must equal the current number of landmarks";
        return EQVIO_ERR_INVALID_ARG;
    }
    stage_reset(f);
    if (n == 0) return EQVIO_OK;
    std::vector<double> soa((size_t)3 * n);
    for (int i = 0; i < n; ++i) {
        soa[i] = p[3 * i];
        soa[n + i] = p[3 * i + 1];
        soa[2 * n + i] = p[3 * i + 2];
    }
    double* h = static_cast<double*>(stage_alloc(f, (size_t)LM_FIELDS * n * sizeof(double)));
    if (!h) return EQVIO_ERR_CUDA;
    std::memcpy(h, soa.data(), soa.size() * sizeof(double));
    for (int i = 0; i < n; ++i) {
        h[F_QW * n + i] = 1.0;
        h[F_QX * n + i] = 0.0;
        h[F_QY * n + i] = 0.0;
        h[F_QZ * n + i] = 0.0;
        h[F_QA * n + i] = 1.0;
    }
    for (int fld = 0; fld < LM_FIELDS; ++fld)
        CUDA_TRY(f, cudaMemcpyAsync(f->lm[f->lmcur] + (size_t)fld * f->cap, h + (size_t)fld * n, n * sizeof(double),
                                    cudaMemcpyHostToDevice, f->stream));
    int rc = upload(f, f->dids[f->lmcur], ids, n);
    if (rc != EQVIO_OK) return rc;
    f->ids.assign(ids, ids + n);
    fill_ll_diag_kernel<<<dim3(cdiv(3 * n, 128), 3 * n), 128, 0, f->stream>>>(f->Sig[f->cur], f->ld, 3 * n, f->st.initialPointVariance,
                                                                              f->st.initialPointDepthVariance);
    LAUNCH_CHECK(f, "fill_ll_diag_kernel");
    CUDA_TRY(f, cudaStreamSynchronize(f->stream));
    return EQVIO_OK;
}

int eqvio_augment_landmark_states(eqvio_filter* f, int n_new, const int* new_ids, int n_provided, const int* provided_ids,
                                  const double* provided_p) {
    ENTER(f);
    if (n_new < 0 || n_provided < 0 || (n_new > 0 && !new_ids) || (n_provided > 0 && (!provided_ids || !provided_p)))
        return EQVIO_ERR_INVALID_ARG;
    stage_reset(f);
    const int N = (int)f->ids.size();
    // id lookups: binary search when the caller's lists are ascending (they come from std::map / getIds()), else sort a copy
    auto make_index = [](int cnt, const int* v, std::vector<std::pair<int, int>>& out) {
        out.resize(cnt);
        for (int j = 0; j < cnt; ++j) out[j] = {v[j], j};
        if (!std::is_sorted(out.begin(), out.end())) std::sort(out.begin(), out.end());
    };
    auto find_in = [](const std::vector<std::pair<int, int>>& idx, int id) -> int {
        auto it = std::lower_bound(idx.begin(), idx.end(), std::make_pair(id, -1));
        return (it != idx.end() && it->first == id) ? it->second : -1;
    };
    std::vector<std::pair<int, int>> want, prov;
    make_index(n_new, new_ids, want);
    make_index(n_provided, provided_ids, prov);
    // one pass over the state: which state ids are wanted (keep) and which wanted ids are already in the state -- no index of
    // the (unsorted) state id list is needed
    std::vector<char> keep(N, 1), inState(n_new, 0);
    for (int i = 0; i < N; ++i) {
        const int j = find_in(want, f->ids[i]);
        if (j < 0)
            keep[i] = 0;
        else
            inState[j] = 1;
    }
    std::vector<int> addIds;
    std::vector<double> addP;
    for (int j = 0; j < n_new; ++j) {
        if (inState[j]) continue;
        bool dup = false;
        for (int a : addIds) dup |= (a == new_ids[j]);
        if (dup) continue;
        const int pj = find_in(prov, new_ids[j]);
        if (pj < 0) {
            f->err = "augment_landmark_states: a new id is missing from the provided state";
            return EQVIO_ERR_INVALID_ARG;
        }
        addIds.push_back(new_ids[j]);
        for (int a = 0; a < 3; ++a) addP.push_back(provided_p[3 * pj + a]);
    }
    if (f->stageTiming) cudaEventRecord(f->augEv[0], f->stream);
    {
        bool identity = addIds.empty();
        for (int i = 0; identity && i < N; ++i) identity = keep[i] != 0;
        if (!identity) f->estValid = false;
    }
    int rc = remove_and_append(f, keep, addIds, addP, f->st.initialPointVariance, -1.0);
    if (rc != EQVIO_OK) return rc;
    if (f->stageTiming) {
        cudaEventRecord(f->augEv[1], f->stream);
        f->augTimed = true;
    }
    return EQVIO_OK;  // asynchronous: the next call on this handle is ordered behind it on the stream
}

int eqvio_process_imu(eqvio_filter* f, double stamp, const double gyr[3], const double acc[3], const double gyr_bias_vel[3],
                      const double acc_bias_vel[3]) {
    ENTER(f);
    if (!gyr || !acc) return EQVIO_ERR_INVALID_ARG;
    if (!f->initialised) {
        int rc = eqvio_initialise_from_imu(f, stamp, gyr, acc);
        if (rc != EQVIO_OK) return rc;
    }
    ImuSample u;
    u.stamp = stamp;
    for (int i = 0; i < 3; ++i) {
        u.v[i] = gyr[i];
        u.v[3 + i] = acc[i];
        u.v[6 + i] = gyr_bias_vel ? gyr_bias_vel[i] : 0.0;
        u.v[9 + i] = acc_bias_vel ? acc_bias_vel[i] : 0.0;
    }
    f->buf.push_back(u);
    return EQVIO_OK;
}

int eqvio_process_imu_rows(eqvio_filter* f, int count, const double* rows) {
    if (!f || count < 0 || (count > 0 && !rows)) return EQVIO_ERR_INVALID_ARG;
    for (int i = 0; i < count; ++i) {
        const double* r = rows + 13 * (size_t)i;
        int rc = eqvio_process_imu(f, r[0], r + 1, r + 4, r + 7, r + 10);
        if (rc != EQVIO_OK) return rc;
    }
    return EQVIO_OK;
}

int eqvio_process_vision(eqvio_filter* f, double stamp, int n, const int* ids, const double* y, const eqvio_camera* cam,
                         int* did_update) {
    ENTER(f);
    if (did_update) *did_update = 0;
    using clk = std::chrono::steady_clock;
    auto us = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    const auto t0 = clk::now();
    int rc = vision_phase_a(f, stamp, n, ids, y, cam);
    const auto t1 = clk::now();
    if (rc == EQVIO_OK) rc = vision_phase_b(f);
    const auto t2 = clk::now();
    f->hostWaitUs = 0;
    int rc2 = vision_phase_c(f, did_update);
    const auto t3 = clk::now();
    f->hostUs[0] += us(t0, t1);
    f->hostUs[1] += us(t1, t2);
    f->hostUs[2] += f->hostWaitUs;
    f->hostUs[3] += us(t2, t3) - f->hostWaitUs;
    ++f->hostCalls;
    return rc != EQVIO_OK ? rc : rc2;
}

int eqvio_get_host_profile(eqvio_filter* f, int reset, double us[4], long long* calls) {
    if (!f || !us) return EQVIO_ERR_INVALID_ARG;
    for (int i = 0; i < 4; ++i) us[i] = f->hostUs[i];
    if (calls) *calls = f->hostCalls;
    if (reset) {
        for (int i = 0; i < 4; ++i) f->hostUs[i] = 0;
        f->hostCalls = 0;
    }
    return EQVIO_OK;
}

int eqvio_replay(eqvio_filter* f, int count, const eqvio_replay_frame* frames, const eqvio_camera* cam, size_t flush_bytes,
                 double* frame_ms, double* est_sensor) {
    if (!f || count < 0 || (count > 0 && !frames) || !cam) return EQVIO_ERR_INVALID_ARG;
    using clk = std::chrono::steady_clock;
    if (cudaSetDevice(f->device) != cudaSuccess) return EQVIO_ERR_CUDA;
    void* scratch = nullptr;
    if (flush_bytes > 0 && cudaMalloc(&scratch, flush_bytes) != cudaSuccess) {
        f->err = "eqvio_replay: cannot allocate the L2 flush scratch";
        return EQVIO_ERR_CUDA;
    }
    int rc = EQVIO_OK;
    std::vector<int> ids;
    std::vector<double> p;
    for (int k = 0; k < count && rc == EQVIO_OK; ++k) {
        const eqvio_replay_frame& fr = frames[k];
        if (scratch) {  // outside the bracket: evict Sigma and everything else from L2
            cudaMemsetAsync(scratch, k & 0xFF, flush_bytes, f->stream);
            cudaStreamSynchronize(f->stream);
        }
        const auto t0 = clk::now();
        for (int i = 0; i < fr.n_imu && rc == EQVIO_OK; ++i) {
            const double* r = fr.imu_rows + 13 * (size_t)i;
            rc = eqvio_process_imu(f, r[0], r + 1, r + 4, r + 7, r + 10);
        }
        if (rc == EQVIO_OK && fr.provided_p) rc = eqvio_augment_landmark_states(f, fr.n, fr.ids, fr.n, fr.ids, fr.provided_p);
        int did = 0;
        if (rc == EQVIO_OK) rc = eqvio_process_vision(f, fr.stamp, fr.n, fr.ids, fr.y, cam, &did);
        if (rc == EQVIO_OK) {
            const int N = eqvio_num_landmarks(f);
            ids.resize(std::max(N, 1));
            p.resize(3 * (size_t)std::max(N, 1));
            double sensor[23];
            int n_out = 0;
            rc = eqvio_get_state_estimate(f, sensor, ids.data(), p.data(), &n_out);
            if (rc == EQVIO_OK && est_sensor) std::memcpy(est_sensor + 23 * (size_t)k, sensor, sizeof(sensor));
        }
        if (frame_ms) frame_ms[k] = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
    }
    if (scratch) cudaFree(scratch);
    return rc;
}

int eqvio_replay_batch(eqvio_filter* const* fs, int count, int n_frames, const eqvio_replay_frame* const* frames, const eqvio_camera* cam,
                       double* frame_ms, double* est_sensor, double* wall_ms) {
    if (!fs || count < 0 || n_frames < 0 || (count > 0 && !frames) || !cam) return EQVIO_ERR_INVALID_ARG;
    std::vector<int> rc(count, EQVIO_OK);
    std::vector<std::thread> th;
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < count; ++k)
        th.emplace_back([&, k] {
            rc[k] = eqvio_replay(fs[k], n_frames, frames[k], cam, 0, frame_ms ? frame_ms + (size_t)k * n_frames : nullptr,
                                 est_sensor ? est_sensor + (size_t)k * n_frames * 23 : nullptr);
        });
    for (auto& t : th) t.join();
    if (wall_ms) *wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (int k = 0; k < count; ++k)
        if (rc[k] != EQVIO_OK) return rc[k];
    return EQVIO_OK;
}

int eqvio_batch_process_vision(eqvio_filter* const* fs, int count, const double* stamps, const int* n, const int* const* ids,
                               const double* const* y, const eqvio_camera* cam, int* did_update) {
    if (!fs || count < 0 || !stamps || !n || !ids || !y || !cam) return EQVIO_ERR_INVALID_ARG;
    int worst = EQVIO_OK;
    std::vector<int> rc(count, EQVIO_OK);
    for (int k = 0; k < count; ++k) {
        if (!fs[k]) return EQVIO_ERR_INVALID_ARG;
        if (cudaSetDevice(fs[k]->device) != cudaSuccess) return EQVIO_ERR_CUDA;
        rc[k] = vision_phase_a(fs[k], stamps[k], n[k], ids[k], y[k], cam);
    }
    for (int k = 0; k < count; ++k) {
        cudaSetDevice(fs[k]->device);
        if (rc[k] == EQVIO_OK) rc[k] = vision_phase_b(fs[k]);
    }
    for (int k = 0; k < count; ++k) {
        cudaSetDevice(fs[k]->device);
        int d = 0;
        int r2 = vision_phase_c(fs[k], &d);
        if (rc[k] == EQVIO_OK) rc[k] = r2;
        if (did_update) did_update[k] = (rc[k] == EQVIO_OK) ? d : 0;
        if (rc[k] != EQVIO_OK && worst == EQVIO_OK) worst = rc[k];
    }
    return worst;
}

double eqvio_get_time(const eqvio_filter* f) { return f ? f->time : -1.0; }
int eqvio_is_initialised(const eqvio_filter* f) { return f && f->initialised ? 1 : 0; }
int eqvio_num_landmarks(const eqvio_filter* f) { return f ? (int)f->ids.size() : 0; }
int eqvio_state_dim(const eqvio_filter* f) { return f ? SENSOR_DIM + 3 * (int)f->ids.size() : 0; }
int eqvio_capacity(const eqvio_filter* f) { return f ? f->cap : 0; }

int eqvio_get_state_estimate(eqvio_filter* f, double sensor[23], int* ids, double* p, int* n_out) {
    ENTER(f);
    const int N = (int)f->ids.size();
    if (f->estValid) {  // produced by the last steady update, state untouched since
        const double* h = reinterpret_cast<const double*>(f->h_out + f->outOffEst);
        if (sensor) std::memcpy(sensor, h, 23 * sizeof(double));
        if (ids && N) std::memcpy(ids, f->ids.data(), N * sizeof(int));
        if (p && N) std::memcpy(p, h + 23, 3 * (size_t)N * sizeof(double));
        if (n_out) *n_out = N;
        return EQVIO_OK;
    }
    stage_reset(f);
    launch_pdl(f, state_estimate_kernel, dim3(cdiv(std::max(N, 1), 128)), dim3(128), (size_t)(0), f->stream, f->lm[f->lmcur], f->cap, N, f->d_xi0s, f->d_Xs[f->xcur], f->d_out, TL_SLOT(f));
    LAUNCH_CHECK(f, "state_estimate_kernel");
    double* h = nullptr;
    int rc = download_async(f, &h, f->d_out, 23 + 3 * (size_t)N);
    if (rc != EQVIO_OK) return rc;
    CUDA_TRY(f, cudaStreamSynchronize(f->stream));
    if (sensor) std::memcpy(sensor, h, 23 * sizeof(double));
    if (ids && N) std::memcpy(ids, f->ids.data(), N * sizeof(int));
    if (p && N) std::memcpy(p, h + 23, 3 * (size_t)N * sizeof(double));
    if (n_out) *n_out = N;
    return EQVIO_OK;
}

int eqvio_get_eqf_state(eqvio_filter* f, double xi0_sensor[23], int* ids, double* xi0_p, double X_group[23], double* X_Q,
                        double* Sigma, int ld) {
    ENTER(f);
    stage_reset(f);
    const int N = (int)f->ids.size();
    const int dim = SENSOR_DIM + 3 * N;
    if (xi0_sensor) std::memcpy(xi0_sensor, f->xi0s, 23 * sizeof(double));
    if (ids && N) std::memcpy(ids, f->ids.data(), N * sizeof(int));
    int rc;
    double *hX = nullptr, *hlm = nullptr;
    if ((rc = download_async(f, &hX, f->d_Xs[f->xcur], 23)) != EQVIO_OK) return rc;
    if (N > 0 && (xi0_p || X_Q)) {
        hlm = static_cast<double*>(stage_alloc(f, (size_t)LM_FIELDS * N * sizeof(double)));
        if (!hlm) return EQVIO_ERR_CUDA;
        for (int fld = 0; fld < LM_FIELDS; ++fld)
            CUDA_TRY(f, cudaMemcpyAsync(hlm + (size_t)fld * N, f->lm[f->lmcur] + (size_t)fld * f->cap, N * sizeof(double),
                                        cudaMemcpyDeviceToHost, f->stream));
    }
    if (Sigma) {
        if (ld < dim) {
            f->err = "get_eqf_state: ld < dim";
            return EQVIO_ERR_INVALID_ARG;
        }
        pack_sigma_kernel<<<dim3(cdiv(dim, 128), dim), 128, 0, f->stream>>>(f->Sig[f->cur], f->ld, dim, f->d_Z, dim);
        LAUNCH_CHECK(f, "pack_sigma_kernel");
        CUDA_TRY(f, cudaMemcpy2DAsync(Sigma, (size_t)ld * sizeof(double), f->d_Z, (size_t)dim * sizeof(double),
                                      (size_t)dim * sizeof(double), dim, cudaMemcpyDeviceToHost, f->stream));
    }
    CUDA_TRY(f, cudaStreamSynchronize(f->stream));
    if (X_group) std::memcpy(X_group, hX, 23 * sizeof(double));
    if (hlm) {
        for (int i = 0; i < N; ++i) {
            if (xi0_p) {
                xi0_p[3 * i] = hlm[F_Q0X * N + i];
                xi0_p[3 * i + 1] = hlm[F_Q0Y * N + i];
                xi0_p[3 * i + 2] = hlm[F_Q0Z * N + i];
            }
            if (X_Q) {
                X_Q[5 * i] = hlm[F_QW * N + i];
                X_Q[5 * i + 1] = hlm[F_QX * N + i];
                X_Q[5 * i + 2] = hlm[F_QY * N + i];
                X_Q[5 * i + 3] = hlm[F_QZ * N + i];
                X_Q[5 * i + 4] = hlm[F_QA * N + i];
            }
        }
    }
    return EQVIO_OK;
}

int eqvio_get_landmark_cov_blocks(eqvio_filter* f, double* blocks) {
    ENTER(f);
    stage_reset(f);
    const int N = (int)f->ids.size();
    if (N == 0) return EQVIO_OK;
    if (!blocks) return EQVIO_ERR_INVALID_ARG;
    cov_blocks_kernel<<<cdiv(N, 128), 128, 0, f->stream>>>(f->Sig[f->cur], f->ld, N, f->d_Z);
    LAUNCH_CHECK(f, "cov_blocks_kernel");
    double* h = nullptr;
    int rc = download_async(f, &h, f->d_Z, 9 * (size_t)N);
    if (rc != EQVIO_OK) return rc;
    CUDA_TRY(f, cudaStreamSynchronize(f->stream));
    std::memcpy(blocks, h, 9 * (size_t)N * sizeof(double));
    return EQVIO_OK;
}

int eqvio_get_feature_predictions(eqvio_filter* f, const eqvio_camera* cam, double stamp, int* ids, double* y, int* n_out) {
    ENTER(f);
    if (n_out) *n_out = 0;
    if (!f->st.useFeaturePredictions) return EQVIO_OK;  // VIOFilter.cpp:247-252: an empty measurement
    if (!cam) return EQVIO_ERR_INVALID_ARG;
    if (cam->model != EQVIO_CAMERA_PINHOLE && cam->model != EQVIO_CAMERA_RADTAN && cam->model != EQVIO_CAMERA_EQUIDISTANT) {
        f->err = "unsupported camera model";
        return EQVIO_ERR_UNSUPPORTED;
    }
    stage_reset(f);
    const int N = (int)f->ids.size();
    if (n_out) *n_out = N;
    if (N == 0) return EQVIO_OK;
    // predictState (VIO_eqf.cpp:139-151): zero-order-hold segments of the buffered IMU samples up to `stamp`
    const int n = (int)f->buf.size();
    std::vector<double> rows((size_t)13 * std::max(n, 1), 0.0);
    for (int i = 0; i < n; ++i) {
        const double t0 = std::max(f->buf[i].stamp, f->time);
        const double t1 = i + 1 < n ? std::min(f->buf[i + 1].stamp, stamp) : stamp;
        rows[13 * i] = std::max(t1 - t0, 0.0);
        for (int k = 0; k < 12; ++k) rows[13 * i + 1 + k] = f->buf[i].v[k];
    }
    int rc;
    double* d_rows = f->d_Z;  // scratch
    if (n > 0 && (rc = upload(f, d_rows, rows.data(), (size_t)13 * n)) != EQVIO_OK) return rc;
    double* d_px = f->d_Z + (size_t)13 * std::max(n, 1);
    predict_kernel<<<cdiv(N, 128), 128, 0, f->stream>>>(f->lm[f->lmcur], f->cap, N, f->d_xi0s, f->d_Xs[f->xcur], d_rows, n,
                                                         to_camera(cam), d_px);
    LAUNCH_CHECK(f, "predict_kernel");
    double* h = nullptr;
    if ((rc = download_async(f, &h, d_px, 2 * (size_t)N)) != EQVIO_OK) return rc;
    CUDA_TRY(f, cudaStreamSynchronize(f->stream));
    if (ids) std::memcpy(ids, f->ids.data(), N * sizeof(int));
    if (y) std::memcpy(y, h, 2 * (size_t)N * sizeof(double));
    return EQVIO_OK;
}

// VIO_eqf::computeNEES (VIO_eqf.cpp:153-170): eps^T Sigma^-1 eps / dim with eps the chart coordinates of the true
// state seen through X^-1.  Sigma^-1 is never formed: Sigma = L L^T by the blocked sweep, z = L^-1 eps, NEES = |z|^2 / dim.
int eqvio_compute_nees(eqvio_filter* f, const double true_sensor[23], int n_true, const int* true_ids, const double* true_p,
                       double* nees) {
    ENTER(f);
    if (!true_sensor || !nees || n_true < 0 || (n_true > 0 && (!true_ids || !true_p))) return EQVIO_ERR_INVALID_ARG;
    stage_reset(f);
    const int N = (int)f->ids.size();
    const int dim = SENSOR_DIM + 3 * N;
    // truncate / reorder the true landmarks to the state order (VIO_eqf.cpp:154-163)
    std::vector<std::pair<int, int>> idx(n_true);
    for (int j = 0; j < n_true; ++j) idx[j] = {true_ids[j], j};
    std::sort(idx.begin(), idx.end());
    std::vector<double> tp((size_t)3 * std::max(N, 1));
    for (int i = 0; i < N; ++i) {
        auto it = std::lower_bound(idx.begin(), idx.end(), std::make_pair(f->ids[i], -1));
        if (it == idx.end() || it->first != f->ids[i]) {
            f->err = "compute_nees: a state landmark is missing from the true state";
            return EQVIO_ERR_INVALID_ARG;
        }
        for (int a = 0; a < 3; ++a) tp[3 * i + a] = true_p[3 * it->second + a];
    }
    int rc;
    const int Mz = dim + 1;
    const int ldz = (Mz + 7) & ~7;
    double* Z = f->d_Z;
    double* d_ts = f->d_Gamma;        // scratch: 23 doubles
    double* d_tp = f->d_uv;           // scratch: 3N doubles (UV_STRIDE * cap available)
    if ((rc = upload(f, d_ts, true_sensor, 23)) != EQVIO_OK) return rc;
    if (N > 0 && (rc = upload(f, d_tp, tp.data(), (size_t)3 * N)) != EQVIO_OK) return rc;
    CUDA_TRY(f, cudaMemsetAsync(f->d_status, 0, sizeof(int), f->stream));
    pack_sigma_kernel<<<dim3(cdiv(dim, 128), dim), 128, 0, f->stream>>>(f->Sig[f->cur], f->ld, dim, Z, ldz);
    LAUNCH_CHECK(f, "pack_sigma_kernel");
    nees_eps_kernel<<<cdiv(std::max(N, 1), 128), 128, 0, f->stream>>>(f->lm[f->lmcur], f->cap, N, f->d_xi0s, f->d_Xs[f->xcur], d_ts, d_tp,
                                                                      f->st.coordinateChoice, Z, ldz, dim);
    LAUNCH_CHECK(f, "nees_eps_kernel");
    for (int k = 0; k < dim; k += NB) {
        const int nbk = std::min(NB, dim - k);
        const int below = Mz - (k + nbk);
        chol_panel_kernel<<<std::max(1, cdiv(below, PANEL_THREADS)), PANEL_THREADS, 0, f->stream>>>(Z, ldz, Mz, k, nbk, f->d_status,
                                                                                                     f->d_Lout);
        LAUNCH_CHECK(f, "chol_panel_kernel");
        if (k + nbk < dim) {
            const int o = k + nbk;
            gemm_nt_sub_kernel<false><<<dim3(cdiv(dim - o, GBN), cdiv(Mz - o, GBM)), 128, 0, f->stream>>>(
                Z + (size_t)o * ldz + o, ldz, Z + (size_t)k * ldz + o, ldz, Z + (size_t)k * ldz + o, ldz, Mz - o, dim - o, nbk);
            LAUNCH_CHECK(f, "gemm_nt_sub_kernel<trail>");
        }
    }
    // z^T is the last row of Z: gather it (stride ldz) with a 2-D copy
    double* hz = static_cast<double*>(stage_alloc(f, (size_t)dim * sizeof(double)));
    int* hst = nullptr;
    if (!hz) return EQVIO_ERR_CUDA;
    CUDA_TRY(f, cudaMemcpy2DAsync(hz, sizeof(double), Z + dim, (size_t)ldz * sizeof(double), sizeof(double), dim,
                                  cudaMemcpyDeviceToHost, f->stream));
    if ((rc = download_async(f, &hst, f->d_status, 1)) != EQVIO_OK) return rc;
    CUDA_TRY(f, cudaStreamSynchronize(f->stream));
    if (*hst & 1) {
        f->err = "compute_nees: Sigma is not positive definite";
        return EQVIO_ERR_NUMERIC;
    }
    double s2 = 0.0;
    for (int k = 0; k < dim; ++k) s2 += hz[k] * hz[k];
    *nees = s2 / dim;
    return EQVIO_OK;
}

int eqvio_get_last_outliers(const eqvio_filter* f, int* ids, int cap, int* n_out) {
    if (!f) return EQVIO_ERR_INVALID_ARG;
    const int k = (int)f->lastOutliers.size();
    if (n_out) *n_out = k;
    if (ids)
        for (int i = 0; i < std::min(k, cap); ++i) ids[i] = f->lastOutliers[i];
    return EQVIO_OK;
}

int eqvio_get_stage_ms(eqvio_filter* f, double ms[3]) {
    if (!f || !ms) return EQVIO_ERR_INVALID_ARG;
    for (int i = 0; i < 3; ++i) ms[i] = f->stageMs[i];
    return EQVIO_OK;
}
int eqvio_enable_stage_timing(eqvio_filter* f, int on) {
    if (!f) return EQVIO_ERR_INVALID_ARG;
    f->stageTiming = on != 0;
    return EQVIO_OK;
}
long long eqvio_get_launch_count(const eqvio_filter* f) { return f ? f->launches : 0; }
int eqvio_get_graph_stats(const eqvio_filter* f, long long* captures, long long* replays) {
    if (!f) return EQVIO_ERR_INVALID_ARG;
    if (captures) *captures = f->graphCaptures;
    if (replays) *replays = f->graphLaunches;
    return EQVIO_OK;
}

int eqvio_enable_kernel_profile(eqvio_filter* f, int on) {
    if (!f) return EQVIO_ERR_INVALID_ARG;
    f->profiling = on != 0;
    return EQVIO_OK;
}
int eqvio_get_kernel_profile(eqvio_filter* f, int reset, double ms[EQVIO_PROF_CLASSES], long long launches[EQVIO_PROF_CLASSES]) {
    if (!f) return EQVIO_ERR_INVALID_ARG;
    for (int i = 0; i < PROF_CLASSES; ++i) {
        if (ms) ms[i] = f->profMs[i];
        if (launches) launches[i] = f->profLaunches[i];
        if (reset) {
            f->profMs[i] = 0;
            f->profLaunches[i] = 0;
        }
    }
    return EQVIO_OK;
}

static void clear_graphs(eqvio_filter* f) {
    for (auto& g : f->graphs)
        if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    f->graphs.clear();
}

int eqvio_set_tuning(eqvio_filter* f, int key, int value) {
    if (!f) return EQVIO_ERR_INVALID_ARG;
    switch (key) {
        case EQVIO_TUNE_CORRECTION:
            if (value < 0 || value > 2) return EQVIO_ERR_INVALID_ARG;
            f->corrMode = value;
            clear_graphs(f);
            return EQVIO_OK;
        case EQVIO_TUNE_DOWNDATE:
            if (value != 0 && value != 1) return EQVIO_ERR_INVALID_ARG;
            f->downdateTC = value;
            clear_graphs(f);
            return EQVIO_OK;
        case EQVIO_TUNE_SPECULATE_NEW:
            f->specNew = value != 0;
            return EQVIO_OK;
        case EQVIO_TUNE_STAGE_S:
            f->stageS = value != 0;
            clear_graphs(f);
            return EQVIO_OK;
        case EQVIO_TUNE_ZERO_COPY:
            f->zeroCopy = value != 0;
            clear_graphs(f);
            return EQVIO_OK;
        case EQVIO_TUNE_PROP_FUSION:
            f->propFusion = value != 0;
            clear_graphs(f);
            return EQVIO_OK;
        case EQVIO_TUNE_FUSE_SMALL:
            f->fuseSmall = value != 0;
            clear_graphs(f);
            return EQVIO_OK;
        case EQVIO_TUNE_PDL:
            f->pdl = value != 0;
            clear_graphs(f);
            return EQVIO_OK;
        case EQVIO_TUNE_FUSE_OBSERVER:
            f->fuseObserver = value != 0;
            clear_graphs(f);
            return EQVIO_OK;
        case EQVIO_TUNE_LOOKAHEAD:
            if (value < 0 || value > 2) return EQVIO_ERR_INVALID_ARG;
            f->lookahead = value;
            clear_graphs(f);
            return EQVIO_OK;
        case EQVIO_TUNE_LAZY_DOWNDATE:
            if (value < 0 || value > 8) return EQVIO_ERR_INVALID_ARG;
            f->lazyMerge = value;
            clear_graphs(f);
            return EQVIO_OK;
        case EQVIO_TUNE_GRAPH:
            f->useGraph = value != 0;
            return EQVIO_OK;
        case EQVIO_TUNE_SPECULATE:
            f->speculate = value != 0;
            return EQVIO_OK;
        case EQVIO_TUNE_CHUNK_LANDMARKS:
            if (value < 1 || value > CH_R / 2) return EQVIO_ERR_INVALID_ARG;
            f->chunkLm = value;
            return EQVIO_OK;
        default:
            return EQVIO_ERR_INVALID_ARG;
    }
}

#ifdef EQVIO_TIMELINE
// out[2 * slot] / out[2 * slot + 1] = first start / last end (globaltimer ns) of the chunk kernels launched since the last
// reset, in launch order (factor, downdate parts ...); returns the number of slots used.  reset re-arms the stamps.
int eqvio_debug_timeline(eqvio_filter* f, unsigned long long* out, int maxSlots, int reset) {
    if (!f) return -1;
    cudaStreamSynchronize(f->stream);
    const int n = std::min(TL_MAX, maxSlots);  // graph replays keep the slots baked in at capture: the caller filters unstamped slots
    if (out && n > 0) cudaMemcpyFromSymbol(out, g_tl, sizeof(unsigned long long) * 2 * n);
    if (reset) {
        std::vector<unsigned long long> init(2 * TL_MAX);
        for (int i = 0; i < TL_MAX; ++i) {
            init[2 * i] = ~0ull;
            init[2 * i + 1] = 0;
        }
        cudaMemcpyToSymbol(g_tl, init.data(), sizeof(unsigned long long) * 2 * TL_MAX);
        f->tlNext = 0;
    }
    return n;
}
#endif

#ifdef EQVIO_TIMELINE
const char* eqvio_debug_timeline_name(eqvio_filter* f, int slot) {
    return (f && slot >= 0 && slot < 512 && f->tlNames[slot]) ? f->tlNames[slot] : "?";
}
#endif

#ifdef EQVIO_CHUNK_TIMING
int eqvio_debug_chunk_timing(long long out[16]) {
    return cudaMemcpyFromSymbol(out, g_chunk_t, sizeof(long long) * 16) == cudaSuccess ? 0 : -2;
}
int eqvio_debug_chunk_fine(long long out[128]) {
    return cudaMemcpyFromSymbol(out, g_chunk_fine, sizeof(long long) * 128) == cudaSuccess ? 0 : -2;
}
int eqvio_debug_bc_gt(unsigned long long out[32]) { return cudaMemcpyFromSymbol(out, g_bc_gt, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -2; }
int eqvio_debug_bc_timing(long long out[16], int warps[BC_S_WARPS * 64]) {
    if (cudaMemcpyFromSymbol(out, g_bc_t, sizeof(long long) * 16) != cudaSuccess) return -2;
    return cudaMemcpyFromSymbol(warps, g_bc_warp, sizeof(int) * BC_S_WARPS * 64) == cudaSuccess ? 0 : -2;
}
#endif

int eqvio_plan_lazy_downdates(int T, int nchunks, const int* band_lo, const int* band_hi, int M, int* out, int max_steps) {
    if (T < 1 || nchunks < 1 || M < 1 || !band_lo || !band_hi || !out || max_steps < nchunks) return EQVIO_ERR_INVALID_ARG;
    std::vector<LazyStep> plan;
    plan_lazy_downdates(T, nchunks, band_lo, band_hi, M, plan);
    for (int c = 0; c < nchunks; ++c) {
        const LazyStep& s = plan[c];
        const int v[8] = {s.blo, s.bhi, s.nBand, s.waitRest, s.hasRest, s.xlo, s.xhi, s.nRest};
        for (int k = 0; k < 8; ++k) out[8 * c + k] = v[k];
    }
    return nchunks;
}

const char* eqvio_build_info(void) { return "eqvio_b200 sm_100a fp64 (CUDA " EQVIO_STR(__CUDACC_VER_MAJOR__) "." EQVIO_STR(__CUDACC_VER_MINOR__) ")"; }

}  // extern "C"
