// multi.cu — tnb_multi_contract_path: the single-process, multi-GPU form of tnb_contract_path.
//
// Who needs it: the reference runs every routine from ONE Julia task (SURVEY §8b "who calls it"), so a `contract(tn; path)`
// that wants the 8 GPUs of a box cannot rely on an external one-process-per-GPU launcher.  This entry point takes the
// leaves on the caller's context (device d0), spawns one host thread + one context per additional device, creates a
// communicator over them (one ncclCommInitRank per thread with a shared id — the multi-threaded equivalent of
// ncclCommInitAll), broadcasts the leaves (they are KBs next to the GiB-sized intermediates), deals the slices round-robin
// (slice s on rank s mod ngpus, hoisted steps redone on every rank), sums the per-GPU accumulators with ONE all-reduce and
// leaves the result in `out` on the caller's device.  The worker contexts and the communicator are cached in the caller's
// context and torn down with it.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "tnb_internal.h"

// comm.cu
int tnb_comm_broadcast_bytes(tnb_ctx* ctx, void* ptr, size_t bytes, int root);

struct tnb_multi {
    int ngpus = 0;
    std::vector<tnb_ctx*> ctx;        // [0] is the caller's context (not owned)
};

static void multi_destroy(tnb_multi* m) {
    if (!m) return;
    for (size_t r = 1; r < m->ctx.size(); r++) tnb_ctx_destroy(m->ctx[r]);    // destroys their communicators too
    delete m;
}

void tnb_multi_release(tnb_ctx* ctx) {
    if (!ctx || !ctx->multi) return;
    multi_destroy((tnb_multi*)ctx->multi);
    ctx->multi = nullptr;
}

// element span [lo, hi] a descriptor can reach
static void span_of(const tnb_tensor* T, int64_t* lo, int64_t* hi) {
    *lo = *hi = T->offset_elems;
    for (int r = 0; r < T->rank; r++) {
        const int64_t s = (T->extent[r] - 1) * T->stride_elems[r];
        if (s > 0) *hi += s; else *lo += s;
    }
}

static int multi_setup(tnb_ctx* ctx0, int ngpus) {
    tnb_multi* m = (tnb_multi*)ctx0->multi;
    if (m && m->ngpus == ngpus) return TNB_OK;
    if (ctx0->comm) return tnb_set_error(ctx0, TNB_EINVAL, "multi_contract_path: the context already belongs to a one-process-per-GPU communicator");
    if (m) { multi_destroy(m); ctx0->multi = nullptr; }
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ngpus > ndev) return tnb_set_error(ctx0, TNB_EINVAL, "multi_contract_path: %d GPUs requested, %d visible", ngpus, ndev);
    m = new tnb_multi();
    m->ngpus = ngpus;
    m->ctx.push_back(ctx0);
    // ranks 1.. take the other devices in ascending order
    for (int d = 0, r = 1; d < ndev && r < ngpus; d++) {
        if (d == ctx0->device) continue;
        tnb_ctx* c = nullptr;
        int rc = tnb_ctx_create(d, &c);
        if (rc) { std::string why = tnb_last_error(nullptr); multi_destroy(m); return tnb_set_error(ctx0, rc, "multi_contract_path: context on device %d: %s", d, why.c_str()); }
        c->c64_mode = ctx0->c64_mode; c->force_generic = ctx0->force_generic; c->gemm_pair = ctx0->gemm_pair;
        m->ctx.push_back(c);
        r++;
    }
    char id[128];
    int rc = tnb_comm_unique_id(id);
    if (rc) { std::string why = tnb_last_error(nullptr); multi_destroy(m); return tnb_set_error(ctx0, rc, "%s", why.c_str()); }
    std::vector<int> rcs(ngpus, 0);
    std::vector<std::thread> th;
    for (int r = 1; r < ngpus; r++) th.emplace_back([&, r] { rcs[r] = tnb_comm_init(m->ctx[r], id, r, ngpus); });
    rcs[0] = tnb_comm_init(ctx0, id, 0, ngpus);
    for (auto& t : th) t.join();
    for (int r = 0; r < ngpus; r++)
        if (rcs[r]) {
            std::string why = m->ctx[r]->err;
            if (!rcs[0]) tnb_comm_destroy(ctx0);
            multi_destroy(m);
            return tnb_set_error(ctx0, rcs[r], "multi_contract_path: rank %d: %s", r, why.c_str());
        }
    ctx0->multi = m;
    return TNB_OK;
}

extern "C" {

int tnb_multi_contract_path(tnb_ctx* ctx0, const tnb_tensor* leaves, int32_t nleaves, const int32_t* steps, int32_t nsteps,
                            const int32_t* sliced_modes, int32_t nsliced, const tnb_tensor* out, int32_t ngpus) {
    if (!ctx0) return TNB_EINVAL;
    if (!leaves || nleaves < 1 || !out || !out->buf) return tnb_set_error(ctx0, TNB_EINVAL, "multi_contract_path: bad arguments");
    if (ngpus < 1) return tnb_set_error(ctx0, TNB_EINVAL, "multi_contract_path: ngpus = %d", ngpus);
    if (ngpus == 1)
        return tnb_contract_path(ctx0, leaves, nleaves, steps, nsteps, sliced_modes, nsliced, 0, 1, INT64_MAX, out);
    // the accumulators are all-reduced in place: the output view must be dense (any index order)
    int64_t olo, ohi, oelems = 1;
    span_of(out, &olo, &ohi);
    for (int r = 0; r < out->rank; r++) oelems *= out->extent[r];
    if (ohi - olo + 1 != oelems) return tnb_set_error(ctx0, TNB_EUNSUPPORTED, "multi_contract_path: the output view must be dense");
    int rc = multi_setup(ctx0, ngpus);
    if (rc) return rc;
    tnb_multi* m = (tnb_multi*)ctx0->multi;
    const size_t esz = tnb_dtype_size(out->dtype);

    std::vector<int> rcs(ngpus, 0);
    auto rank_main = [&](int r) {
        tnb_ctx* c = m->ctx[r];
        cudaSetDevice(c->device);
        std::vector<tnb_tensor> L(leaves, leaves + nleaves);
        std::vector<tnb_buf*> owned;
        tnb_tensor O = *out;
        tnb_buf* obuf = nullptr;
        int rc = TNB_OK;
        // leaves: same descriptors over a private copy of each leaf's span, filled by a broadcast from rank 0
        for (int i = 0; i < nleaves && !rc; i++) {
            int64_t lo, hi;
            span_of(&leaves[i], &lo, &hi);
            const size_t lsz = tnb_dtype_size(leaves[i].dtype);
            const size_t bytes = (size_t)(hi - lo + 1) * lsz;
            void* ptr;
            if (r == 0) {
                ptr = (char*)leaves[i].buf->ptr + (size_t)lo * lsz;
            } else {
                tnb_buf* b = nullptr;
                if ((rc = tnb_alloc(c, bytes, &b))) break;
                owned.push_back(b);
                L[i].buf = b;
                L[i].offset_elems = leaves[i].offset_elems - lo;
                ptr = b->ptr;
            }
            rc = tnb_comm_broadcast_bytes(c, ptr, bytes, 0);
        }
        if (!rc && r != 0) {
            if (!(rc = tnb_alloc(c, (size_t)oelems * esz, &obuf))) { O.buf = obuf; O.offset_elems = out->offset_elems - olo; }
        }
        tnb_plan* P = nullptr;
        if (!rc) rc = tnb_plan_create(c, L.data(), nleaves, steps, nsteps, sliced_modes, nsliced, &O, &P);
        if (!rc) rc = tnb_memset_zero(c, O.buf, (size_t)(O.offset_elems + (olo - out->offset_elems)) * esz, (size_t)oelems * esz);
        if (!rc) rc = tnb_plan_execute(c, P, r, ngpus, INT64_MAX, 1);
        // every rank must enter the collective, also after a local failure elsewhere would deadlock the others: a failed
        // rank still calls it on whatever its accumulator holds, and the error is reported after the join
        int rc2 = TNB_OK;
        if (O.buf) rc2 = tnb_comm_allreduce_sum(c, O.buf, (size_t)(O.offset_elems + (olo - out->offset_elems)) * esz, oelems, out->dtype);
        if (!rc) rc = rc2;
        cudaStreamSynchronize(c->stream);
        if (P) tnb_plan_destroy(c, P);
        for (tnb_buf* b : owned) tnb_free(c, b);
        if (obuf) tnb_free(c, obuf);
        rcs[r] = rc;
    };
    std::vector<std::thread> th;
    for (int r = 1; r < ngpus; r++) th.emplace_back(rank_main, r);
    rank_main(0);
    for (auto& t : th) t.join();
    cudaSetDevice(ctx0->device);
    for (int r = 0; r < ngpus; r++)
        if (rcs[r]) {
            if (r != 0) { std::string why = m->ctx[r]->err; return tnb_set_error(ctx0, rcs[r], "multi_contract_path: rank %d: %s", r, why.c_str()); }
            return rcs[r];
        }
    return TNB_OK;
}

}  // extern "C"
