// host_entry.inl -- HOST-pointer conveniences of the C ABI: H2D, the device entry point, D2H,
// one synchronisation at the end.  Inputs in page-locked memory are DMA'd as they are; large inputs in ordinary
// (pageable) memory -- what a numpy caller holds -- go through the workspace's pinned ring, filled by copy threads
// (staging.h); small ones are left to the driver's own staging.
namespace {

// dst (device) <- src (host), in the order of `stream`
int h2d(pcu_b200_workspace* ws, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    if (ws->opts.host_staging != 2 && (ws->opts.host_staging == 1 || HostStager::wants(src, bytes))) {
        if (!ws->stager) ws->stager = new (std::nothrow) HostStager();
        if (ws->stager) {
            const cudaError_t e = ws->stager->copy(dst, src, bytes, stream);
            if (e != cudaSuccess) return fail(PCU_B200_CUDA_ERROR, "staged host-to-device copy failed: %s", cudaGetErrorString(e));
            return PCU_B200_OK;
        }
    }
    PCU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return PCU_B200_OK;
}


template <typename T>
int knn_host(pcu_b200_workspace* ws, const T* query, long long n, const T* dataset, long long m, int k, int squared,
             T* out_dist, long long* out_idx, long long* out_n_tied, pcu_b200_cloud* prepared = nullptr) {
    if (!ws) return fail(PCU_B200_INVALID_ARGUMENT, "null workspace");
    if (k <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "Invalid value for k (%d) must be greater than 0.", k);
    if (prepared != nullptr) { dataset = query; m = 1; }   // placeholders for the checks and the staging layout: nothing of it is copied
    PCU_TRY(check_cloud_args<T>(query, n, dataset, m));
    if (!out_dist || !out_idx) return fail(PCU_B200_INVALID_ARGUMENT, "null output pointer");
    PCU_ON_DEVICE(ws);
    Carver measure(nullptr);
    auto carve = [&](Carver& cv, T*& dq, T*& dd, T*& od, long long*& oi, long long*& nt) {
        dq = cv.take<T>((size_t)3 * n);
        dd = cv.take<T>((size_t)3 * m);
        od = cv.take<T>((size_t)n * k);
        oi = cv.take<long long>((size_t)n * k);
        nt = cv.take<long long>(1);
    };
    T *dq, *dd, *od; long long *oi, *nt;
    carve(measure, dq, dd, od, oi, nt);
    PCU_TRY(ensure_io(ws, measure.off, ws->own_stream));
    Carver cv(ws->io);
    carve(cv, dq, dd, od, oi, nt);
    cudaStream_t st = ws->own_stream;
    mark(ws, 9, st);
    PCU_TRY(h2d(ws, dq, query, sizeof(T) * 3 * n, st));
    if (prepared == nullptr) PCU_TRY(h2d(ws, dd, dataset, sizeof(T) * 3 * m, st));
    PCU_TRY(knn_device<T>(ws, dq, n, dd, m, k, squared, od, oi, nt, st, prepared));
    PCU_CUDA(cudaMemcpyAsync(out_dist, od, sizeof(T) * n * k, cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaMemcpyAsync(out_idx, oi, sizeof(long long) * n * k, cudaMemcpyDeviceToHost, st));
    long long tied = 0;
    unsigned overflows = 0;
    PCU_CUDA(cudaMemcpyAsync(&tied, nt, sizeof(long long), cudaMemcpyDeviceToHost, st));
    mark(ws, 10, st);
    if (ws->replay_overflows)
        PCU_CUDA(cudaMemcpyAsync(&overflows, ws->replay_overflows, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaStreamSynchronize(st));
    if (out_n_tied) *out_n_tied = tied;
    if (tied > 0 && overflows > 0)
        return fail(PCU_B200_INTERNAL, "tie replay: the reference kd-tree of this dataset is deeper than the %d-frame walk "
                    "stack (%u walks affected); neighbour order among exactly equidistant points is not guaranteed", kKdStack, overflows);
    return PCU_B200_OK;
}

// Device-side results of one fused call, contiguous so that ONE copy brings them to the pinned slot.
template <typename T>
struct StatsResult {
    pcu_b200_nn_stats stats[2];
    T value;
};

template <typename T>
int stats_host(pcu_b200_workspace* ws, const T* a, long long n, const T* b, long long m, bool both,
               pcu_b200_nn_stats* out_stats, T* out_value, const pcu_b200_cloud* prepared = nullptr) {
    if (!ws) return fail(PCU_B200_INVALID_ARGUMENT, "null workspace");
    if (prepared != nullptr) { b = a; m = 1; }   // placeholders for the checks and the staging layout: nothing of b is copied
    PCU_TRY(check_cloud_args<T>(a, n, b, m));
    if (!out_stats) return fail(PCU_B200_INVALID_ARGUMENT, "null output pointer");
    PCU_ON_DEVICE(ws);
    Carver measure(nullptr);
    auto carve = [&](Carver& cv, T*& da, T*& db, StatsResult<T>*& dr) {
        da = cv.take<T>((size_t)3 * n);
        db = cv.take<T>((size_t)3 * m);
        dr = cv.take<StatsResult<T>>(1);
    };
    T *da, *db; StatsResult<T>* dr;
    carve(measure, da, db, dr);
    cudaStream_t st = ws->own_stream;
    PCU_TRY(ensure_io(ws, measure.off, st));
    Carver cv(ws->io);
    carve(cv, da, db, dr);
    // the copies run on their own stream: the first cloud is binned while the second is still in flight.
    // (the staging block may have just been re-allocated in the order of `st`: the copies wait for that)
    PCU_CUDA(cudaEventRecord(ws->arrived[0], st));
    PCU_CUDA(cudaStreamWaitEvent(ws->copy_stream, ws->arrived[0], 0));
    mark(ws, 9, st);
    PCU_TRY(h2d(ws, da, a, sizeof(T) * 3 * n, ws->copy_stream));
    PCU_CUDA(cudaEventRecord(ws->arrived[0], ws->copy_stream));
    // the second cloud's copy is issued from inside the launch sequence, after the first cloud's passes have been
    // enqueued: a staged (pageable) copy keeps this thread busy while the GPU already bins the first cloud
    const std::function<int()> second_copy = [&]() -> int {
        if (prepared == nullptr) PCU_TRY(h2d(ws, db, b, sizeof(T) * 3 * m, ws->copy_stream));
        PCU_CUDA(cudaEventRecord(ws->arrived[1], ws->copy_stream));
        return PCU_B200_OK;
    };
    PCU_TRY(stats_device<T>(ws, da, n, db, m, both, dr->stats, both ? &dr->value : nullptr, st, ws->arrived, prepared, &second_copy));
    StatsResult<T>* hr = reinterpret_cast<StatsResult<T>*>(ws->host_slot);
    auto fetch = [&]() -> int {
        PCU_CUDA(cudaMemcpyAsync(hr, dr, sizeof(StatsResult<T>), cudaMemcpyDeviceToHost, st));
        mark(ws, 10, st);
        PCU_CUDA(cudaStreamSynchronize(st));
        return PCU_B200_OK;
    };
    PCU_TRY(fetch());
    // Hausdorff witnesses whose neighbour was decided by tie order: replay with the reference's tree
    if (ws->opts.disable_tie_replay != 1) {
        for (int s = 0; s < (both ? 2 : 1); ++s) {
            if (!hr->stats[s].witness_tied) continue;
            const T* second = prepared ? (const T*)prepared->raw : db;
            const long long msecond = prepared ? prepared->n : m;
            const T* qs = s == 0 ? da : second;
            const T* dsrc = s == 0 ? second : da;
            PCU_TRY(resolve_witness_device<T>(ws, qs, s == 0 ? n : msecond, dsrc, s == 0 ? msecond : n, dr->stats + s, st));
            PCU_TRY(fetch());
        }
    }
    for (int s = 0; s < (both ? 2 : 1); ++s) out_stats[s] = hr->stats[s];
    if (both && out_value) *out_value = hr->value;
    return PCU_B200_OK;
}

template <typename T>
int batched_chamfer_host(pcu_b200_workspace* ws, const T* x, const T* y, long long batch, long long n, long long m,
                         T* out_per_pair, double* out_sum) {
    if (!ws) return fail(PCU_B200_INVALID_ARGUMENT, "null workspace");
    if (batch <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "batch must be positive (got %lld)", batch);
    PCU_TRY(check_cloud_args<T>(x, n, y, m));
    if (!out_per_pair) return fail(PCU_B200_INVALID_ARGUMENT, "null output pointer");
    PCU_ON_DEVICE(ws);
    Carver measure(nullptr);
    auto carve = [&](Carver& cv, T*& dx, T*& dy, T*& dv, double*& dsum) {
        dx = cv.take<T>((size_t)3 * n * batch);
        dy = cv.take<T>((size_t)3 * m * batch);
        dv = cv.take<T>((size_t)batch);
        dsum = cv.take<double>(1);
    };
    T *dx, *dy, *dv; double* dsum;
    carve(measure, dx, dy, dv, dsum);
    PCU_TRY(ensure_io(ws, measure.off, ws->own_stream));
    Carver cv(ws->io);
    carve(cv, dx, dy, dv, dsum);
    cudaStream_t st = ws->own_stream;
    PCU_TRY(h2d(ws, dx, x, sizeof(T) * 3 * n * batch, st));
    PCU_TRY(h2d(ws, dy, y, sizeof(T) * 3 * m * batch, st));
    PCU_TRY(batched_chamfer_device<T>(ws, dx, dy, batch, n, m, dv, out_sum ? dsum : nullptr, st));
    PCU_CUDA(cudaMemcpyAsync(out_per_pair, dv, sizeof(T) * batch, cudaMemcpyDeviceToHost, st));
    if (out_sum) PCU_CUDA(cudaMemcpyAsync(out_sum, dsum, sizeof(double), cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaStreamSynchronize(st));
    return PCU_B200_OK;
}

// Host wrapper shared by the two normal estimators: stage points (and view directions), run `device_call` on the
// workspace's stream, fetch the count, then exactly the kept rows.
template <typename T, typename DeviceCall>
int normals_host(pcu_b200_workspace* ws, const T* points, long long n, const T* view_dirs, long long* out_idx, T* out_normals,
                 long long* out_count, DeviceCall&& device_call) {
    if (!ws) return fail(PCU_B200_INVALID_ARGUMENT, "null workspace");
    if (!points || n <= 0)
        return fail(PCU_B200_INVALID_ARGUMENT, "Invalid point set with zero elements: points must have shape (n, 3) (got %lld rows)", n);
    if (!out_idx || !out_normals || !out_count) return fail(PCU_B200_INVALID_ARGUMENT, "null output pointer");
    PCU_ON_DEVICE(ws);
    Carver measure(nullptr);
    auto carve = [&](Carver& cv, T*& dp, T*& dv, long long*& di, T*& dn, long long*& dc) {
        dp = cv.take<T>((size_t)3 * n);
        dv = cv.take<T>(view_dirs ? (size_t)3 * n : 1);
        di = cv.take<long long>((size_t)n);
        dn = cv.take<T>((size_t)3 * n);
        dc = cv.take<long long>(1);
    };
    T *dp, *dv, *dn; long long *di, *dc;
    carve(measure, dp, dv, di, dn, dc);
    cudaStream_t st = ws->own_stream;
    PCU_TRY(ensure_io(ws, measure.off, st));
    Carver cv(ws->io);
    carve(cv, dp, dv, di, dn, dc);
    PCU_TRY(h2d(ws, dp, points, sizeof(T) * 3 * n, st));
    if (view_dirs) PCU_TRY(h2d(ws, dv, view_dirs, sizeof(T) * 3 * n, st));
    PCU_TRY(device_call(dp, view_dirs ? dv : (const T*)nullptr, di, dn, dc, st));
    long long* hc = reinterpret_cast<long long*>(ws->host_slot);
    PCU_CUDA(cudaMemcpyAsync(hc, dc, sizeof(long long), cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaStreamSynchronize(st));
    const long long kept = *hc;
    if (kept > 0) {
        PCU_CUDA(cudaMemcpyAsync(out_idx, di, sizeof(long long) * kept, cudaMemcpyDeviceToHost, st));
        PCU_CUDA(cudaMemcpyAsync(out_normals, dn, sizeof(T) * 3 * kept, cudaMemcpyDeviceToHost, st));
        PCU_CUDA(cudaStreamSynchronize(st));
    }
    *out_count = kept;
    return PCU_B200_OK;
}

template <typename T>
int normals_knn_host(pcu_b200_workspace* ws, const T* points, long long n, const T* view_dirs, int k,
                     double drop_angle_threshold, long long* out_idx, T* out_normals, long long* out_count) {
    if (k <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "Invalid number of neighbors (%d) must be greater than 0.", k);
    return normals_host<T>(ws, points, n, view_dirs, out_idx, out_normals, out_count,
                           [&](const T* dp, const T* dv, long long* di, T* dn, long long* dc, cudaStream_t st) {
                               return normals_knn_device<T>(ws, dp, n, dv, k, drop_angle_threshold, di, dn, dc, st);
                           });
}

template <typename T>
int normals_ball_host(pcu_b200_workspace* ws, const T* points, long long n, const T* view_dirs, const pcu_b200_ball_options* o,
                      long long* out_idx, T* out_normals, long long* out_count) {
    return normals_host<T>(ws, points, n, view_dirs, out_idx, out_normals, out_count,
                           [&](const T* dp, const T* dv, long long* di, T* dn, long long* dc, cudaStream_t st) {
                               return normals_ball_device<T>(ws, dp, n, dv, o, di, dn, dc, st);
                           });
}

// voxel-grid down-sampling on HOST arrays: stage, run, fetch the row count, then exactly the rows that exist
template <typename T>
int voxel_downsample_host(pcu_b200_workspace* ws, const T* points, long long n, const void* attrib, int attrib_cols, int attrib_is_f64,
                          const double size[3], const double min_bound[3], const double max_bound[3], int min_points, T* out_points,
                          void* out_attrib, int32_t* out_counts, long long* out_rows) {
    if (!ws) return fail(PCU_B200_INVALID_ARGUMENT, "null workspace");
    if (!points || n <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "points must be a non-empty (n, 3) array");
    if (!out_points || !out_rows || attrib_cols < 0 || (attrib_cols > 0 && (!attrib || !out_attrib)))
        return fail(PCU_B200_INVALID_ARGUMENT, "null pointer");
    PCU_ON_DEVICE(ws);
    const size_t abytes = (size_t)(attrib_is_f64 ? 8 : 4) * (size_t)attrib_cols;      // per row
    Carver measure(nullptr);
    auto carve = [&](Carver& cv, T*& dp, unsigned char*& da, T*& op, unsigned char*& oa, int*& oc, long long*& orows) {
        dp = cv.take<T>((size_t)3 * n);
        da = cv.take<unsigned char>(abytes * (size_t)n + 1);
        op = cv.take<T>((size_t)3 * n);
        oa = cv.take<unsigned char>(abytes * (size_t)n + 1);
        oc = cv.take<int>((size_t)n);
        orows = cv.take<long long>(1);
    };
    T *dp, *op; unsigned char *da, *oa; int* oc; long long* orows;
    carve(measure, dp, da, op, oa, oc, orows);
    cudaStream_t st = ws->own_stream;
    PCU_TRY(ensure_io(ws, measure.off, st));
    Carver cv(ws->io);
    carve(cv, dp, da, op, oa, oc, orows);
    PCU_TRY(h2d(ws, dp, points, sizeof(T) * 3 * n, st));
    if (attrib_cols > 0) PCU_TRY(h2d(ws, da, attrib, abytes * (size_t)n, st));
    PCU_TRY(voxel_downsample_dispatch<T>(ws, dp, n, attrib_cols > 0 ? da : nullptr, attrib_cols, attrib_is_f64, size, min_bound, max_bound,
                                         min_points, op, attrib_cols > 0 ? oa : nullptr, out_counts ? oc : nullptr, orows, st));
    long long* hr = reinterpret_cast<long long*>(ws->host_slot);
    PCU_CUDA(cudaMemcpyAsync(hr, orows, sizeof(long long), cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaStreamSynchronize(st));
    const long long rows = *hr;
    if (rows > 0) {
        PCU_CUDA(cudaMemcpyAsync(out_points, op, sizeof(T) * 3 * rows, cudaMemcpyDeviceToHost, st));
        if (attrib_cols > 0) PCU_CUDA(cudaMemcpyAsync(out_attrib, oa, abytes * (size_t)rows, cudaMemcpyDeviceToHost, st));
        if (out_counts) PCU_CUDA(cudaMemcpyAsync(out_counts, oc, sizeof(int) * rows, cudaMemcpyDeviceToHost, st));
        PCU_CUDA(cudaStreamSynchronize(st));
    }
    *out_rows = rows;
    return PCU_B200_OK;
}

// duplicate removal on HOST arrays (faces may be null)
template <typename T>
int deduplicate_host(pcu_b200_workspace* ws, const T* points, long long n, double epsilon, const void* faces, long long nf, int cols,
                     int faces_are_i64, T* out_points, int32_t* out_svi, int32_t* out_svj, void* out_faces, long long* out_counts) {
    if (!ws) return fail(PCU_B200_INVALID_ARGUMENT, "null workspace");
    if (!points || n <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "points must be a non-empty (n, 3) array");
    if (!out_points || !out_svi || !out_svj || !out_counts) return fail(PCU_B200_INVALID_ARGUMENT, "null output pointer");
    if (faces != nullptr && (nf < 0 || cols <= 0 || cols > 16 || (nf > 0 && !out_faces)))
        return fail(PCU_B200_INVALID_ARGUMENT, "faces must be an (m, c) array with 1 <= c <= 16");
    PCU_ON_DEVICE(ws);
    const bool with_faces = faces != nullptr && nf > 0;
    const size_t fbytes = with_faces ? (size_t)(faces_are_i64 ? 8 : 4) * (size_t)cols * (size_t)nf : 0;
    Carver measure(nullptr);
    auto carve = [&](Carver& cv, T*& dp, unsigned char*& df, T*& op, int*& svi, int*& svj, unsigned char*& of, long long*& oc) {
        dp = cv.take<T>((size_t)3 * n);
        df = cv.take<unsigned char>(fbytes + 1);
        op = cv.take<T>((size_t)3 * n);
        svi = cv.take<int>((size_t)n);
        svj = cv.take<int>((size_t)n);
        of = cv.take<unsigned char>(fbytes + 1);
        oc = cv.take<long long>(3);
    };
    T *dp, *op; unsigned char *df, *of; int *svi, *svj; long long* oc;
    carve(measure, dp, df, op, svi, svj, of, oc);
    cudaStream_t st = ws->own_stream;
    PCU_TRY(ensure_io(ws, measure.off, st));
    Carver cv(ws->io);
    carve(cv, dp, df, op, svi, svj, of, oc);
    PCU_TRY(h2d(ws, dp, points, sizeof(T) * 3 * n, st));
    if (with_faces) PCU_TRY(h2d(ws, df, faces, fbytes, st));
    PCU_TRY(dedup_dispatch<T>(ws, dp, n, epsilon, with_faces ? df : nullptr, with_faces ? nf : 0, cols, faces_are_i64, op, svi, svj,
                              with_faces ? of : nullptr, oc, st));
    long long* hc = reinterpret_cast<long long*>(ws->host_slot);
    PCU_CUDA(cudaMemcpyAsync(hc, oc, 3 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaStreamSynchronize(st));
    const long long unique = hc[0], kept = hc[1];
    PCU_CUDA(cudaMemcpyAsync(out_points, op, sizeof(T) * 3 * unique, cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaMemcpyAsync(out_svi, svi, sizeof(int) * unique, cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaMemcpyAsync(out_svj, svj, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    if (with_faces && kept > 0)
        PCU_CUDA(cudaMemcpyAsync(out_faces, of, (size_t)(faces_are_i64 ? 8 : 4) * (size_t)cols * (size_t)kept, cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 3; ++i) out_counts[i] = hc[i];
    return PCU_B200_OK;
}

// Prepares a cloud from HOST points: staged through the workspace, binned into the handle's own block.
template <typename T>
int cloud_prepare_host(pcu_b200_workspace* ws, const T* points, long long n, pcu_b200_cloud** out, int knn_k = 1, int leaf = 0) {
    if (!ws || !out) return fail(PCU_B200_INVALID_ARGUMENT, "null argument");
    PCU_TRY(check_cloud_args<T>(points, n, points, 1));
    PCU_ON_DEVICE(ws);
    cudaStream_t st = ws->own_stream;
    PCU_TRY(ensure_io(ws, align_up(sizeof(T) * 3 * (size_t)n), st));
    PCU_TRY(h2d(ws, ws->io, points, sizeof(T) * 3 * n, st));
    PCU_TRY(cloud_prepare_device<T>(ws, reinterpret_cast<const T*>(ws->io), n, out, st, knn_k, leaf));
    PCU_CUDA(cudaStreamSynchronize(st));
    return PCU_B200_OK;
}

// Morton entry points on HOST arrays: up to two inputs, one output, copies either side of one device call.
template <typename Call>
int morton_host(pcu_b200_workspace* ws, const void* in_a, size_t bytes_a, const void* in_b, size_t bytes_b, void* out,
                size_t bytes_out, Call&& call) {
    if (!ws) return fail(PCU_B200_INVALID_ARGUMENT, "null workspace");
    if (!in_a || bytes_a == 0 || (bytes_b != 0 && !in_b)) return fail(PCU_B200_INVALID_ARGUMENT, "codes / pts must be non-empty arrays but got an empty array");
    if (!out) return fail(PCU_B200_INVALID_ARGUMENT, "null output pointer");
    PCU_ON_DEVICE(ws);
    cudaStream_t st = ws->own_stream;
    Carver measure(nullptr);
    measure.take<unsigned char>(bytes_a); measure.take<unsigned char>(bytes_b ? bytes_b : 1); measure.take<unsigned char>(bytes_out);
    PCU_TRY(ensure_io(ws, measure.off, st));
    Carver cv(ws->io);
    unsigned char* da = cv.take<unsigned char>(bytes_a);
    unsigned char* db = cv.take<unsigned char>(bytes_b ? bytes_b : 1);
    unsigned char* dout = cv.take<unsigned char>(bytes_out);
    PCU_TRY(h2d(ws, da, in_a, bytes_a, st));
    if (bytes_b) PCU_TRY(h2d(ws, db, in_b, bytes_b, st));
    PCU_TRY(call(da, db, dout, st));
    PCU_CUDA(cudaMemcpyAsync(out, dout, bytes_out, cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaStreamSynchronize(st));
    return PCU_B200_OK;
}

template <typename T>
int debug_kd_tree(pcu_b200_workspace* ws, const T* points, long long m, int leaf, int32_t* order, long long node_cap,
                  int32_t* feat, T* div_lo, T* div_hi, int32_t* first, int32_t* last, int32_t* kid0, int32_t* kid1,
                  int64_t* out_nodes) {
    if (!ws || !points || m <= 0 || leaf <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "bad argument");
    PCU_ON_DEVICE(ws);
    KdReplayBuffers<T> rb;
    T* dpts;
    Carver measure(nullptr);
    rb.carve(measure, m);
    dpts = measure.take<T>((size_t)3 * m);
    PCU_TRY(ensure_arena(ws, measure.off, ws->own_stream));
    Carver cv(ws->arena);
    rb.carve(cv, m);
    dpts = cv.take<T>((size_t)3 * m);
    cudaStream_t st = ws->own_stream;
    PCU_TRY(h2d(ws, dpts, points, sizeof(T) * 3 * m, st));
    const int rs = build_kd_replica<T>(rb, dpts, m, leaf, nullptr, KdPrune<T>{}, st, g_launches);
    if (rs != PCU_B200_OK) return fail(rs, "kd replica build failed: %s", cudaGetErrorString(cudaGetLastError()));
    KdCounters hc;
    PCU_CUDA(cudaMemcpyAsync(&hc, rb.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaMemcpyAsync(order, rb.order, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
    PCU_CUDA(cudaStreamSynchronize(st));
    std::vector<KdNode<T>> nodes((size_t)hc.n_nodes);
    PCU_CUDA(cudaMemcpy(nodes.data(), rb.nodes, sizeof(KdNode<T>) * nodes.size(), cudaMemcpyDeviceToHost));
    for (long long i = 0; i < hc.n_nodes && i < node_cap; ++i) {
        feat[i] = nodes[i].feat; div_lo[i] = nodes[i].div_lo; div_hi[i] = nodes[i].div_hi;
        first[i] = nodes[i].first; last[i] = nodes[i].last; kid0[i] = nodes[i].kid0; kid1[i] = nodes[i].kid1;
    }
    *out_nodes = hc.n_nodes;
    return PCU_B200_OK;
}

}  // namespace

extern "C" {

int pcu_b200_knn_host_f32(pcu_b200_workspace* ws, const float* query, int64_t n, const float* dataset, int64_t m,
                          int k, int squared, float* out_dist, int64_t* out_idx, int64_t* out_n_tied) {
    return knn_host<float>(ws, query, n, dataset, m, k, squared, out_dist, (long long*)out_idx, (long long*)out_n_tied);
}
int pcu_b200_knn_host_f64(pcu_b200_workspace* ws, const double* query, int64_t n, const double* dataset, int64_t m,
                          int k, int squared, double* out_dist, int64_t* out_idx, int64_t* out_n_tied) {
    return knn_host<double>(ws, query, n, dataset, m, k, squared, out_dist, (long long*)out_idx, (long long*)out_n_tied);
}
int pcu_b200_nn_stats_host_f32(pcu_b200_workspace* ws, const float* query, int64_t n, const float* dataset,
                               int64_t m, pcu_b200_nn_stats* out_stats) {
    return stats_host<float>(ws, query, n, dataset, m, false, out_stats, nullptr);
}
int pcu_b200_nn_stats_host_f64(pcu_b200_workspace* ws, const double* query, int64_t n, const double* dataset,
                               int64_t m, pcu_b200_nn_stats* out_stats) {
    return stats_host<double>(ws, query, n, dataset, m, false, out_stats, nullptr);
}
int pcu_b200_chamfer_host_f32(pcu_b200_workspace* ws, const float* x, int64_t n, const float* y, int64_t m,
                              pcu_b200_nn_stats* out_stats, float* out_value) {
    return stats_host<float>(ws, x, n, y, m, true, out_stats, out_value);
}
int pcu_b200_chamfer_host_f64(pcu_b200_workspace* ws, const double* x, int64_t n, const double* y, int64_t m,
                              pcu_b200_nn_stats* out_stats, double* out_value) {
    return stats_host<double>(ws, x, n, y, m, true, out_stats, out_value);
}

int pcu_b200_debug_kd_tree_f32(pcu_b200_workspace* ws, const float* points, int64_t m, int max_points_per_leaf,
                               int32_t* order, int64_t node_cap, int32_t* feat, float* div_lo, float* div_hi,
                               int32_t* first, int32_t* last, int32_t* kid0, int32_t* kid1, int64_t* out_nodes) {
    return debug_kd_tree<float>(ws, points, m, max_points_per_leaf, order, node_cap, feat, div_lo, div_hi, first, last,
                                kid0, kid1, out_nodes);
}
int pcu_b200_debug_kd_tree_f64(pcu_b200_workspace* ws, const double* points, int64_t m, int max_points_per_leaf,
                               int32_t* order, int64_t node_cap, int32_t* feat, double* div_lo, double* div_hi,
                               int32_t* first, int32_t* last, int32_t* kid0, int32_t* kid1, int64_t* out_nodes) {
    return debug_kd_tree<double>(ws, points, m, max_points_per_leaf, order, node_cap, feat, div_lo, div_hi, first, last,
                                 kid0, kid1, out_nodes);
}
int pcu_b200_voxel_downsample_host_f32(pcu_b200_workspace* ws, const float* points, int64_t n, const void* attrib, int attrib_cols,
                                       int attrib_is_f64, const double voxel_size[3], const double min_bound[3], const double max_bound[3],
                                       int min_points_per_voxel, float* out_points, void* out_attrib, int32_t* out_counts, int64_t* out_rows) {
    return voxel_downsample_host<float>(ws, points, n, attrib, attrib_cols, attrib_is_f64, voxel_size, min_bound, max_bound,
                                        min_points_per_voxel, out_points, out_attrib, out_counts, (long long*)out_rows);
}
int pcu_b200_voxel_downsample_host_f64(pcu_b200_workspace* ws, const double* points, int64_t n, const void* attrib, int attrib_cols,
                                       int attrib_is_f64, const double voxel_size[3], const double min_bound[3], const double max_bound[3],
                                       int min_points_per_voxel, double* out_points, void* out_attrib, int32_t* out_counts, int64_t* out_rows) {
    return voxel_downsample_host<double>(ws, points, n, attrib, attrib_cols, attrib_is_f64, voxel_size, min_bound, max_bound,
                                         min_points_per_voxel, out_points, out_attrib, out_counts, (long long*)out_rows);
}
int pcu_b200_deduplicate_host_f32(pcu_b200_workspace* ws, const float* points, int64_t n, double epsilon, const void* faces, int64_t n_faces,
                                  int face_cols, int faces_are_i64, float* out_points, int32_t* out_svi, int32_t* out_svj, void* out_faces,
                                  int64_t* out_counts) {
    return deduplicate_host<float>(ws, points, n, epsilon, faces, n_faces, face_cols, faces_are_i64, out_points, out_svi, out_svj, out_faces,
                                   (long long*)out_counts);
}
int pcu_b200_deduplicate_host_f64(pcu_b200_workspace* ws, const double* points, int64_t n, double epsilon, const void* faces, int64_t n_faces,
                                  int face_cols, int faces_are_i64, double* out_points, int32_t* out_svi, int32_t* out_svj, void* out_faces,
                                  int64_t* out_counts) {
    return deduplicate_host<double>(ws, points, n, epsilon, faces, n_faces, face_cols, faces_are_i64, out_points, out_svi, out_svj, out_faces,
                                    (long long*)out_counts);
}
int pcu_b200_cloud_prepare_knn_host_f32(pcu_b200_workspace* ws, const float* points, int64_t n, int k, int max_points_per_leaf,
                                        pcu_b200_cloud** out_cloud) {
    return cloud_prepare_host<float>(ws, points, n, out_cloud, k, max_points_per_leaf > 0 ? max_points_per_leaf : 10);
}
int pcu_b200_cloud_prepare_knn_host_f64(pcu_b200_workspace* ws, const double* points, int64_t n, int k, int max_points_per_leaf,
                                        pcu_b200_cloud** out_cloud) {
    return cloud_prepare_host<double>(ws, points, n, out_cloud, k, max_points_per_leaf > 0 ? max_points_per_leaf : 10);
}
int pcu_b200_knn_prepared_host_f32(pcu_b200_workspace* ws, const float* query, int64_t n, pcu_b200_cloud* dataset, int k, int squared,
                                   float* out_dist, int64_t* out_idx, int64_t* out_n_tied) {
    if (!dataset) return fail(PCU_B200_INVALID_ARGUMENT, "null prepared cloud");
    return knn_host<float>(ws, query, n, nullptr, 0, k, squared, out_dist, (long long*)out_idx, (long long*)out_n_tied, dataset);
}
int pcu_b200_knn_prepared_host_f64(pcu_b200_workspace* ws, const double* query, int64_t n, pcu_b200_cloud* dataset, int k, int squared,
                                   double* out_dist, int64_t* out_idx, int64_t* out_n_tied) {
    if (!dataset) return fail(PCU_B200_INVALID_ARGUMENT, "null prepared cloud");
    return knn_host<double>(ws, query, n, nullptr, 0, k, squared, out_dist, (long long*)out_idx, (long long*)out_n_tied, dataset);
}
int pcu_b200_cloud_prepare_host_f32(pcu_b200_workspace* ws, const float* points, int64_t n, pcu_b200_cloud** out_cloud) {
    return cloud_prepare_host<float>(ws, points, n, out_cloud);
}
int pcu_b200_cloud_prepare_host_f64(pcu_b200_workspace* ws, const double* points, int64_t n, pcu_b200_cloud** out_cloud) {
    return cloud_prepare_host<double>(ws, points, n, out_cloud);
}
int pcu_b200_chamfer_prepared_host_f32(pcu_b200_workspace* ws, const float* x, int64_t n, const pcu_b200_cloud* y,
                                       pcu_b200_nn_stats* out_stats, float* out_value) {
    if (!y) return fail(PCU_B200_INVALID_ARGUMENT, "null prepared cloud");
    return stats_host<float>(ws, x, n, nullptr, 0, true, out_stats, out_value, y);
}
int pcu_b200_chamfer_prepared_host_f64(pcu_b200_workspace* ws, const double* x, int64_t n, const pcu_b200_cloud* y,
                                       pcu_b200_nn_stats* out_stats, double* out_value) {
    if (!y) return fail(PCU_B200_INVALID_ARGUMENT, "null prepared cloud");
    return stats_host<double>(ws, x, n, nullptr, 0, true, out_stats, out_value, y);
}
int pcu_b200_nn_stats_prepared_host_f32(pcu_b200_workspace* ws, const float* query, int64_t n, const pcu_b200_cloud* dataset,
                                        pcu_b200_nn_stats* out_stats) {
    if (!dataset) return fail(PCU_B200_INVALID_ARGUMENT, "null prepared cloud");
    return stats_host<float>(ws, query, n, nullptr, 0, false, out_stats, nullptr, dataset);
}
int pcu_b200_nn_stats_prepared_host_f64(pcu_b200_workspace* ws, const double* query, int64_t n, const pcu_b200_cloud* dataset,
                                        pcu_b200_nn_stats* out_stats) {
    if (!dataset) return fail(PCU_B200_INVALID_ARGUMENT, "null prepared cloud");
    return stats_host<double>(ws, query, n, nullptr, 0, false, out_stats, nullptr, dataset);
}

int pcu_b200_morton_encode_host_i32(pcu_b200_workspace* ws, const int32_t* pts, int64_t n, uint64_t* out_codes) {
    if (n <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "pts must be an array of shape [n, 3] but got an empty array");
    return morton_host(ws, pts, sizeof(int32_t) * 3 * n, nullptr, 0, out_codes, sizeof(uint64_t) * n,
                       [&](unsigned char* a, unsigned char*, unsigned char* o, cudaStream_t st) {
                           return morton_encode_device<int32_t>(ws, (const int32_t*)a, n, (unsigned long long*)o, st); });
}
int pcu_b200_morton_encode_host_i64(pcu_b200_workspace* ws, const int64_t* pts, int64_t n, uint64_t* out_codes) {
    if (n <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "pts must be an array of shape [n, 3] but got an empty array");
    return morton_host(ws, pts, sizeof(int64_t) * 3 * n, nullptr, 0, out_codes, sizeof(uint64_t) * n,
                       [&](unsigned char* a, unsigned char*, unsigned char* o, cudaStream_t st) {
                           return morton_encode_device<long long>(ws, (const long long*)a, n, (unsigned long long*)o, st); });
}
int pcu_b200_morton_decode_host(pcu_b200_workspace* ws, const uint64_t* codes, int64_t n, int32_t* out_pts) {
    if (n <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "codes must be an array of shape [n] but got an empty array");
    return morton_host(ws, codes, sizeof(uint64_t) * n, nullptr, 0, out_pts, sizeof(int32_t) * 3 * n,
                       [&](unsigned char* a, unsigned char*, unsigned char* o, cudaStream_t st) {
                           return morton_decode_device(ws, (const unsigned long long*)a, n, (int*)o, st); });
}
int pcu_b200_morton_add_host(pcu_b200_workspace* ws, const uint64_t* a, const uint64_t* b, int64_t n, uint64_t* out) {
    if (n <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "codes_1 must be an array of shape [n,] but got an empty array");
    return morton_host(ws, a, sizeof(uint64_t) * n, b, sizeof(uint64_t) * n, out, sizeof(uint64_t) * n,
                       [&](unsigned char* da, unsigned char* db, unsigned char* o, cudaStream_t st) {
                           return morton_addsub_device(ws, (const unsigned long long*)da, (const unsigned long long*)db, n, 0, (unsigned long long*)o, st); });
}
int pcu_b200_morton_subtract_host(pcu_b200_workspace* ws, const uint64_t* a, const uint64_t* b, int64_t n, uint64_t* out) {
    if (n <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "codes_1 must be an array of shape [n,] but got an empty array");
    return morton_host(ws, a, sizeof(uint64_t) * n, b, sizeof(uint64_t) * n, out, sizeof(uint64_t) * n,
                       [&](unsigned char* da, unsigned char* db, unsigned char* o, cudaStream_t st) {
                           return morton_addsub_device(ws, (const unsigned long long*)da, (const unsigned long long*)db, n, 1, (unsigned long long*)o, st); });
}
int pcu_b200_morton_knn_host(pcu_b200_workspace* ws, const uint64_t* codes, int64_t n, const uint64_t* qcodes, int64_t m, int k,
                             int sort_dist, int64_t* out_idx) {
    if (k <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "k must be greater than 0");
    if (n <= 0 || m <= 0) return fail(PCU_B200_INVALID_ARGUMENT, "codes must be an array of shape [n] but got an empty array");
    if ((int64_t)k > n) return fail(PCU_B200_INVALID_ARGUMENT, "k (%d) exceeds the number of codes (%lld): clamp it first (morton.cpp:351)", k, (long long)n);
    return morton_host(ws, codes, sizeof(uint64_t) * n, qcodes, sizeof(uint64_t) * m, out_idx, sizeof(int64_t) * m * k,
                       [&](unsigned char* da, unsigned char* db, unsigned char* o, cudaStream_t st) {
                           return morton_knn_device(ws, (const unsigned long long*)da, n, (const unsigned long long*)db, m, k, sort_dist, (long long*)o, st); });
}

int pcu_b200_normals_knn_host_f32(pcu_b200_workspace* ws, const float* points, int64_t n, const float* view_dirs, int k,
                                  double drop_angle_threshold, int64_t* out_idx, float* out_normals, int64_t* out_count) {
    return normals_knn_host<float>(ws, points, n, view_dirs, k, drop_angle_threshold, (long long*)out_idx, out_normals, (long long*)out_count);
}
int pcu_b200_normals_knn_host_f64(pcu_b200_workspace* ws, const double* points, int64_t n, const double* view_dirs, int k,
                                  double drop_angle_threshold, int64_t* out_idx, double* out_normals, int64_t* out_count) {
    return normals_knn_host<double>(ws, points, n, view_dirs, k, drop_angle_threshold, (long long*)out_idx, out_normals, (long long*)out_count);
}
int pcu_b200_normals_ball_host_f32(pcu_b200_workspace* ws, const float* points, int64_t n, const float* view_dirs,
                                   const pcu_b200_ball_options* options, int64_t* out_idx, float* out_normals, int64_t* out_count) {
    return normals_ball_host<float>(ws, points, n, view_dirs, options, (long long*)out_idx, out_normals, (long long*)out_count);
}
int pcu_b200_normals_ball_host_f64(pcu_b200_workspace* ws, const double* points, int64_t n, const double* view_dirs,
                                   const pcu_b200_ball_options* options, int64_t* out_idx, double* out_normals, int64_t* out_count) {
    return normals_ball_host<double>(ws, points, n, view_dirs, options, (long long*)out_idx, out_normals, (long long*)out_count);
}
int pcu_b200_batched_chamfer_host_f32(pcu_b200_workspace* ws, const float* x, const float* y, int64_t batch,
                                      int64_t n, int64_t m, float* out_per_pair, double* out_sum) {
    return batched_chamfer_host<float>(ws, x, y, batch, n, m, out_per_pair, out_sum);
}

}  // extern "C"
