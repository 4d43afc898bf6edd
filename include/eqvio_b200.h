/*
 * eqvio_b200 -- C ABI of the B200-native EqF vision-update path.
 *
 * This is the drop-in boundary for the hot path of pvangoor/eqvio: everything
 * the reference's `VIOFilter` (include/eqvio/VIOFilter.h:36-192) does between
 * `processIMUData` / `processVisionData` and `stateEstimate`, executed on one
 * B200 with the Riccati matrix Sigma resident in HBM.  The reference has no FFI
 * of its own; each entry point below names the C++ member it replaces so that a
 * re-bodied `VIOFilter` can forward to it one-to-one.
 *
 * Conventions
 *   - plain C types only; all pointers are HOST pointers unless a name ends in
 *     `_dev`; arrays are caller-owned, sizes are explicit.
 *   - return value: 0 = EQVIO_OK.  Conditions on which the reference returns
 *     silently (filter not initialised, time not advanced, empty measurement;
 *     src/VIOFilter.cpp:135-136,198-199,223-224) are NOT errors: the call
 *     returns EQVIO_OK and sets *did_update = 0.  Negative values are errors;
 *     `eqvio_last_error` gives the message.  Nothing throws across the ABI.
 *   - one handle = one filter = one CUDA stream; a handle is not thread-safe,
 *     distinct handles are independent (Monte-Carlo replicas).
 *   - sensor state, flat `sensor[23]`:
 *       [0,6) inputBias (gyr, acc) | [6,10) pose quaternion (w,x,y,z) |
 *       [10,13) pose position | [13,16) body velocity |
 *       [16,20) camera-offset quaternion | [20,23) camera-offset position
 *     group element X, flat `group[23]`:
 *       [0,6) beta | [6,10) A quaternion | [10,13) A translation | [13,16) w |
 *       [16,20) B quaternion | [20,23) B translation
 *   - landmark order is the reference's state order (insertion order of
 *     xi0.cameraLandmarks / X.id); measurements are given in ascending id like
 *     the reference's std::map (src/mathematical/VisionMeasurement.cpp:24-28).
 *   - Sigma is exported column-major, dim x dim, dim = 21 + 3 N, in the
 *     reference's state-vector order (coordinateSuite/euclid.cpp:103-109).
 */
#ifndef EQVIO_B200_H
#define EQVIO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define EQVIO_OK 0
#define EQVIO_ERR_INVALID_ARG (-1)
#define EQVIO_ERR_CUDA (-2)
#define EQVIO_ERR_NUMERIC (-3)     /* NaN / non-SPD innovation covariance detected on device */
#define EQVIO_ERR_CAPACITY (-4)    /* more landmarks than the handle was created for */
#define EQVIO_ERR_UNSUPPORTED (-5) /* something this build has no CUDA path for (an unknown coordinateChoice, a camera model other
                                      than pinhole / radtan / equidistant) */

#define EQVIO_COORD_EUCLIDEAN 0
#define EQVIO_COORD_INVDEPTH 1
#define EQVIO_COORD_NORMAL 2 /* normal.cpp: A = M A_euclid M^-1 with the numerically differentiated, block-diagonal chart change M, applied block by block */

#define EQVIO_CAMERA_PINHOLE 0
#define EQVIO_CAMERA_RADTAN 1
#define EQVIO_CAMERA_EQUIDISTANT 2 /* Kannala-Brandt fisheye, dist[0..3] (GIFT EquidistantCamera) */

typedef struct eqvio_filter eqvio_filter;

/* POD mirror of VIOFilter::Settings (include/eqvio/VIOFilterSettings.h:58-99);
 * field names and defaults are the reference's. */
typedef struct eqvio_settings {
    double biasOmegaProcessVariance, biasAccelProcessVariance, attitudeProcessVariance, positionProcessVariance,
        velocityProcessVariance, cameraAttitudeProcessVariance, cameraPositionProcessVariance, pointProcessVariance;
    double velGyrNoise, velAccNoise, velGyrBiasWalk, velAccBiasWalk;
    double measurementNoise, outlierThresholdAbs, outlierThresholdProb, featureRetention;
    double initialAttitudeVariance, initialPositionVariance, initialVelocityVariance, initialCameraAttitudeVariance,
        initialCameraPositionVariance, initialPointVariance, initialPointDepthVariance, initialBiasOmegaVariance,
        initialBiasAccelVariance, initialSceneDepth;
    int useDiscreteInnovationLift, useDiscreteVelocityLift, useDiscreteStateMatrix, fastRiccati, useMedianDepth,
        useFeaturePredictions, useEquivariantOutput, removeLostLandmarks;
    int coordinateChoice;   /* EQVIO_COORD_* */
    double cameraOffset[7]; /* quaternion (w,x,y,z), position -- Settings::cameraOffset */
} eqvio_settings;

/* Flattened GIFT::GICamera (external/GIFT/GIFT/include/GIFT/camera/GICamera.h:29-71).
 * For RADTAN the caller supplies inv_dist, i.e. StandardCamera::invDist; use
 * eqvio_camera_fit_inverse_distortion to recompute it (the member is protected in GIFT). */
typedef struct eqvio_camera {
    int model; /* EQVIO_CAMERA_* */
    int width, height;
    int ndist; /* number of valid entries of dist: 0, 2, 4 or 5 (inv_dist always holds the five entries of StandardCamera::invDist) */
    double fx, fy, cx, cy;
    double dist[5];
    double inv_dist[5];
} eqvio_camera;

/* --- settings / camera helpers ------------------------------------------------ */
/* VIOFilter::Settings() defaults (VIOFilterSettings.h:59-98). */
void eqvio_settings_default(eqvio_settings* s);
/* StandardCamera::computeInverseDistortion (external/GIFT/GIFT/src/camera/StandardCamera.cpp:113-145). */
int eqvio_camera_fit_inverse_distortion(eqvio_camera* cam);

/* --- lifetime ------------------------------------------------------------------- */
/* VIOFilter(const Settings&) (src/VIOFilter.cpp:31-41).  `capacity` = the largest
 * landmark count the handle will ever hold (Sigma is allocated capacity-padded in
 * HBM once).  `stream_or_null`: a cudaStream_t to run on, or NULL for a private one. */
int eqvio_create(const eqvio_settings* s, int device, int capacity, void* stream_or_null, eqvio_filter** out);
/* VIOFilter(const VIOState& xi0, const Settings&, const double& time) (src/VIOFilter.cpp:43-56). */
int eqvio_create_from_state(const eqvio_settings* s, int device, int capacity, void* stream_or_null,
                            const double sensor[23], int n, const int* ids, const double* p /* 3n, xyz per landmark */,
                            double time, eqvio_filter** out);
void eqvio_destroy(eqvio_filter* f);
const char* eqvio_last_error(const eqvio_filter* f); /* f may be NULL: last create error */

/* --- state setters ---------------------------------------------------------------- */
int eqvio_initialise_from_imu(eqvio_filter* f, double stamp, const double gyr[3],
                              const double acc[3]); /* initialiseFromIMUData, VIOFilter.cpp:65-78 */
int eqvio_set_state(eqvio_filter* f, const double sensor[23], int n, const int* ids,
                    const double* p); /* setState, VIOFilter.cpp:80-92 */
int eqvio_set_landmarks(eqvio_filter* f, int n, const int* ids,
                        const double* p); /* setLandmarks, VIOFilter.cpp:94-110 */
/* augmentLandmarkStates (VIOFilter.cpp:112-132): keep only landmarks whose id is in
 * new_ids, then append the ids that are new with the position found in the provided state. */
int eqvio_augment_landmark_states(eqvio_filter* f, int n_new, const int* new_ids, int n_provided,
                                  const int* provided_ids, const double* provided_p);

/* --- input --------------------------------------------------------------------------- */
/* processIMUData (VIOFilter.cpp:58-63): buffers the sample; first sample initialises attitude. */
int eqvio_process_imu(eqvio_filter* f, double stamp, const double gyr[3], const double acc[3],
                      const double gyr_bias_vel[3] /* may be NULL = 0 */, const double acc_bias_vel[3] /* may be NULL */);
/* `count` processIMUData calls in one: rows of 13 doubles (stamp, gyr3, acc3, gyrBiasVel3, accBiasVel3). */
int eqvio_process_imu_rows(eqvio_filter* f, int count, const double* rows);
/* processVisionData (VIOFilter.cpp:194-241): propagate to `stamp`, prune lost landmarks, gate
 * outliers, add new landmarks, EqF correction, drop invalid landmarks.  ids ascending. */
int eqvio_process_vision(eqvio_filter* f, double stamp, int n, const int* ids, const double* y /* 2n pixels */,
                         const eqvio_camera* cam, int* did_update /* may be NULL */);
/* Host-side replay of recorded frames through the entry points above, the way the reference's main loop drives VIOFilter
 * (main_sim.cpp:136-148): per frame eqvio_process_imu (one call per row), eqvio_augment_landmark_states, eqvio_process_vision,
 * eqvio_get_state_estimate -- all on HOST buffers.  This is the C++ host loop of the end-to-end measurement: frame_ms[k]
 * (may be NULL) = host wall time of frame k, which ends synchronised with the state estimate on the host; est_sensor
 * (count x 23, may be NULL) receives the sensor part of every estimate.  flush_bytes > 0 writes that many bytes of device
 * scratch between frames, outside the timed bracket (L2 flush).  Stops at the first error and returns it. */
typedef struct eqvio_replay_frame {
    double stamp;
    int n_imu;
    const double* imu_rows; /* n_imu x 13: stamp, gyr3, acc3, gyrBiasVel3, accBiasVel3 */
    int n;
    const int* ids;            /* n measured ids, ascending */
    const double* y;           /* 2n pixels */
    const double* provided_p;  /* n x 3 landmark positions for augmentLandmarkStates (same ids), or NULL = skip the augment call */
} eqvio_replay_frame;
int eqvio_replay(eqvio_filter* f, int count, const eqvio_replay_frame* frames, const eqvio_camera* cam, size_t flush_bytes,
                 double* frame_ms, double* est_sensor);
/* eqvio_replay for `count` independent filters at once, one host thread per filter (Monte-Carlo replicas: their kernels overlap
 * on the GPU through the filters' own streams, their host-side planning on the cores).  frames[k] points to n_frames frames of
 * filter k; frame_ms / est_sensor as in eqvio_replay with filter k's block at offset k * n_frames (may be NULL);
 * wall_ms (may be NULL) = wall time of the whole batch.  Returns the first error. */
int eqvio_replay_batch(eqvio_filter* const* fs, int count, int n_frames, const eqvio_replay_frame* const* frames, const eqvio_camera* cam,
                       double* frame_ms, double* est_sensor, double* wall_ms);
/* Same update for `count` independent filters at once (Monte-Carlo replicas on one GPU):
 * kernels of different filters overlap on their streams.  Arrays are indexed per filter. */
int eqvio_batch_process_vision(eqvio_filter* const* fs, int count, const double* stamps, const int* n,
                               const int* const* ids, const double* const* y, const eqvio_camera* cam,
                               int* did_update /* count entries or NULL */);

/* --- output ------------------------------------------------------------------------------ */
double eqvio_get_time(const eqvio_filter* f);       /* getTime, VIOFilter.cpp:256 */
int eqvio_is_initialised(const eqvio_filter* f);    /* isInitialised, VIOFilter.h:170 */
int eqvio_num_landmarks(const eqvio_filter* f);     /* filterState.X.id.size() */
int eqvio_state_dim(const eqvio_filter* f);         /* xi0.Dim() = 21 + 3N */
int eqvio_capacity(const eqvio_filter* f);
/* stateEstimate (VIOFilter.cpp:243): xi_hat = stateGroupAction(X, xi0).  ids/p sized >= N. */
int eqvio_get_state_estimate(eqvio_filter* f, double sensor[23], int* ids, double* p, int* n_out);
/* viewEqFState (VIOFilter.cpp:245): xi0, X and Sigma.  Any output pointer may be NULL.
 * X_Q holds 5 doubles per landmark: quaternion (w,x,y,z), scale a.  Sigma is dim x dim, column-major, ld >= dim. */
int eqvio_get_eqf_state(eqvio_filter* f, double xi0_sensor[23], int* ids, double* xi0_p, double X_group[23],
                        double* X_Q, double* Sigma, int ld);
/* VIO_eqf::getLandmarkCovById for every landmark: 9 doubles (column-major 3x3) per landmark
 * (VIO_eqf.cpp:188-194; what VIOWriter::writeConsistency reads, src/VIOWriter.cpp:162-222). */
int eqvio_get_landmark_cov_blocks(eqvio_filter* f, double* blocks /* 9N */);
/* getFeaturePredictions (VIOFilter.cpp:247-252): predictState (VIO_eqf.cpp:139-151) over the buffered IMU
 * samples up to `stamp`, projected through `cam`; ids / y sized >= N, state order.  n_out = 0 unless
 * Settings::useFeaturePredictions. */
int eqvio_get_feature_predictions(eqvio_filter* f, const eqvio_camera* cam, double stamp, int* ids, double* y,
                                  int* n_out);
/* VIO_eqf::computeNEES (VIO_eqf.cpp:153-170), what eqvio_sim prints every frame (main_sim.cpp:146-148): the true
 * state (sensor[23] + landmarks by id, any order, must contain every state landmark) -> NEES.  Solved on the device
 * with a Cholesky sweep of Sigma instead of the reference's dense inverse. */
int eqvio_compute_nees(eqvio_filter* f, const double true_sensor[23], int n_true, const int* true_ids, const double* true_p,
                       double* nees);
/* ids removed as outliers by the last eqvio_process_vision (VIOFilter.cpp:304-364), in removal order. */
int eqvio_get_last_outliers(const eqvio_filter* f, int* ids, int cap, int* n_out);

/* --- measurement hooks (bench / tests only) ------------------------------------------------- */
/* Device time in ms of the stages of the last eqvio_process_vision, measured with CUDA events on
 * the handle's stream: [0] propagation, [1] preprocessing (gate + compaction, plus the device time of
 * eqvio_augment_landmark_states calls made since the previous update), [2] correction -- the
 * reference's LoopTimer labels (src/VIOFilter.cpp:196-236). */
int eqvio_get_stage_ms(eqvio_filter* f, double ms[3]);
/* Enable per-stage event timing (adds two event records per stage); off by default. */
int eqvio_enable_stage_timing(eqvio_filter* f, int on);
/* Number of kernel launches issued by this handle since creation. */
long long eqvio_get_launch_count(const eqvio_filter* f);
/* CUDA graphs captured (instantiated) and replayed so far: a steady frame shape is captured once and replayed afterwards. */
int eqvio_get_graph_stats(const eqvio_filter* f, long long* captures, long long* replays);
/* Per-kernel-class device time, measured with CUDA events recorded on the handle's stream around
 * every launch of the class (adds two event records per launch; off by default, not meant to be on
 * while whole-update throughput is timed).  Classes: */
#define EQVIO_PROF_PROP_LL 0 /* Riccati propagation, landmark-landmark block (HBM-bound) */
#define EQVIO_PROF_PANEL 1   /* chunk factorisation + solve (batch mode: panel of the Cholesky sweep) */
#define EQVIO_PROF_TRAIL 2   /* batch mode only: trailing update of the sweep (FP64 tensor-core GEMM) */
#define EQVIO_PROF_SYRK 3    /* Sigma downdate Sigma -= Y^T Y (FP64 tensor cores, DMMA) */
#define EQVIO_PROF_BC_DIAG 4  /* block sweep: diagonal step (look-ahead products + 64 x 64 factorization), the critical chain */
#define EQVIO_PROF_BC_PANEL 5 /* block sweep: panels P = T L^-T (DMMA block substitution) */
#define EQVIO_PROF_BC_TRAIL 6 /* block sweep: trailing tiles of S / W and the Sigma downdate (DMMA) */
#define EQVIO_PROF_CLASSES 7
int eqvio_enable_kernel_profile(eqvio_filter* f, int on);
/* Accumulated ms and launch counts per class since the last reset. */
int eqvio_get_kernel_profile(eqvio_filter* f, int reset, double ms[EQVIO_PROF_CLASSES],
                             long long launches[EQVIO_PROF_CLASSES]);
/* Host-side time of eqvio_process_vision since the last reset, microseconds summed over `calls` calls:
 * us[0] phase A (id matching, frame staging, enqueue / graph launch), us[1] phase B (host decisions of a non-steady
 * frame), us[2] waiting for the device, us[3] the rest of phase C (status check, bookkeeping). */
int eqvio_get_host_profile(eqvio_filter* f, int reset, double us[4], long long* calls);
/* Evaluation-order knobs of the correction (results agree to rounding; tests run both).
 *   EQVIO_TUNE_CORRECTION: 0 = sequential landmark chunks exploiting C's block sparsity,
 *                          1 = batch form: Cholesky sweep over [S; W^T; ytilde^T], then Sigma -= Y^T Y,
 *                          2 (default) = block Cholesky sweep with look-ahead (64-row blocks; the chain of diagonal factorizations on
 *                              one stream, panels / trailing tiles / Sigma downdate beside it) for up to 768 measurement rows,
 *                              sequential chunks beyond.
 *   EQVIO_TUNE_CHUNK_LANDMARKS: landmarks per chunk in mode 0 (1..32, default 32).
 *   EQVIO_TUNE_SPECULATE: 1 (default) = when a frame brings no new ids, launch the correction before the
 *                          gate scalars reach the host, guarded on the device by a "gate tripped" flag, and
 *                          redo it through the exact host decision path if the flag came back set; 0 = always
 *                          wait for the gate.  Same results either way. */
#define EQVIO_TUNE_CORRECTION 0
#define EQVIO_TUNE_CHUNK_LANDMARKS 1
#define EQVIO_TUNE_SPECULATE 2
/*   EQVIO_TUNE_GRAPH: 1 (default) = replay the launch sequence of a steady frame (no landmark enters or leaves)
 *                      as a cached CUDA graph; 0 = issue the launches one by one. */
#define EQVIO_TUNE_GRAPH 3
/*   EQVIO_TUNE_DOWNDATE: 0 (default) = Sigma -= Y^T Y in fp64 on the FP64 tensor pipe (DMMA);
 *                         1 = BASELINE configs[2]: tcgen05 tensor cores, Y split into three bf16 terms (24-bit
 *                         mantissa), fp32 accumulation in TMEM, Sigma itself stays fp64.  Parity with the fp64 path
 *                         is ~1e-7 per update instead of ~1e-15. */
#define EQVIO_TUNE_DOWNDATE 5
/*   EQVIO_TUNE_LOOKAHEAD: 2 (default) = automatic: 1 when the tile grid of Sigma spans more than one wave (>= 24 tile
 *                         rows, N >= ~500), else 0.  1 = each chunk's downdate is split into the tiles the NEXT chunk gathers from
 *                         (its landmarks' tile rows / columns; launched at once) and the remaining lower tiles, which run
 *                         on a second stream beside the next chunk's factor kernel; 0 = one downdate launch per chunk,
 *                         strictly in order.  Same arithmetic per tile either way (bit-identical results). */
#define EQVIO_TUNE_LOOKAHEAD 6
/*   EQVIO_TUNE_FUSE_OBSERVER: 1 (default) = integrateObserverState's serial sensor chain and per-landmark chain run as one
 *                         software-pipelined kernel; 0 = two kernels.  Same arithmetic (bit-identical results). */
#define EQVIO_TUNE_FUSE_OBSERVER 7
/*   EQVIO_TUNE_PDL: 1 (default) = the chunk kernels are launched with programmatic dependent launch allowed (the next
 *                         grid is scheduled while its predecessor drains and blocks in griddepcontrol.wait); 0 = plain. */
#define EQVIO_TUNE_PDL 8
/*   EQVIO_TUNE_FUSE_SMALL: 1 (default) = in a steady update the gate and the measurement rows (C*, ytilde) run as one launch
 *                         and the innovation lift also emits the state estimate; 0 = four separate kernels.  Same arithmetic. */
#define EQVIO_TUNE_FUSE_SMALL 10
/*   EQVIO_TUNE_SPECULATE_NEW: 1 (default) = frames that bring new ids speculate as well: the new landmarks' positions (bearing x
 *                         median scene depth, VIOFilter.cpp:258-278,366-380) are computed on the device from the gate kernel's depths,
 *                         so the update needs no host round trip; if a gate trips they are dropped and added again through the exact
 *                         host path; 0 = wait for the gate scalars whenever a frame brings new ids. */
#define EQVIO_TUNE_SPECULATE_NEW 11
/*   EQVIO_TUNE_STAGE_S: 1 (default) = the chunk factor kernel fetches Sigma[L_c, L_c] as ONE 2-D TMA tensor copy (96 x 96 box of a
 *     CUtensorMap over the covariance) into shared memory when the chunk's landmarks are consecutive in the state; 0 = every tile
 *     owner gathers its 36 entries itself (also the fall-back for non-consecutive chunks).  Same entries, bit-identical results. */
#define EQVIO_TUNE_STAGE_S 12
/*   EQVIO_TUNE_ZERO_COPY: 1 (default) = the per-frame input block and the result block of a steady update move through a copy
 *     KERNEL over host-mapped pinned memory; 0 = cudaMemcpyAsync (memcpy nodes on a copy engine inside the replayed graph). */
#define EQVIO_TUNE_ZERO_COPY 13
/*   EQVIO_TUNE_PROP_FUSION: 1 (default) = in the fast Riccati step the landmark-landmark kernel builds the rank-27 factors of its own
 *     tile rows / columns and the sensor-landmark strip runs beside it on another stream (chain: prologue + rows -> ll);
 *     0 = prologue + rows -> strip (writes the factors) -> ll.  Same expressions, bit-identical results. */
#define EQVIO_TUNE_PROP_FUSION 14
/*   EQVIO_TUNE_LAZY_DOWNDATE: M = 1 (default): in the look-ahead form of the sequential chunks (more than 768 measurement rows) the
 *     panels Y_c of every chunk are kept, a per-tile word says through which chunk a tile of Sigma is current, the deferred launch
 *     of every M-th chunk leaves out the tile rows / columns the next M + 1 chunks gather from, and the urgent launch of a chunk brings
 *     its tiles up to date from wherever they stand (K = 64 .. 64 (M + 1) rows of Y per visit) -- so the chain factor -> urgent tiles
 *     -> factor only waits for a deferred launch issued M + 1 chunks earlier;  0 = every chunk downdates every lower tile (K = 64)
 *     and the urgent launch of chunk c waits for the deferred one of chunk c - 1.  Per tile the same products in the same order:
 *     bit-identical results for every M. */
#define EQVIO_TUNE_LAZY_DOWNDATE 15
int eqvio_set_tuning(eqvio_filter* f, int key, int value);
/* Launch plan of the lazy trailing updates (EQVIO_TUNE_LAZY_DOWNDATE = M) for T x T tiles of Sigma and nchunks landmark chunks;
 * band_lo[c] .. band_hi[c] (c < nchunks - 1) = the tile rows / columns chunk c + 1 gathers from.  Host logic only (no GPU
 * needed): out[8 c ..] = band lo, band hi, urgent tiles, chunk whose deferred launch the urgent launch waits for (-1: none), deferred
 * launch issued (0 / 1), its left-out tile range lo, hi, its tile count.  Returns nchunks, or EQVIO_ERR_INVALID_ARG.  This is the
 * plan eqvio_process_vision executes; tests/test_lazy_plan.py checks its invariants (no tile shared by launches that may overlap,
 * every tile current at the end). */
int eqvio_plan_lazy_downdates(int T, int nchunks, const int* band_lo, const int* band_hi, int M, int* out, int max_steps);
/* Version / build info string (arch the kernels were compiled for). */
const char* eqvio_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* EQVIO_B200_H */
