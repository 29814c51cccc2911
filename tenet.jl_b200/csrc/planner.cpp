// planner.cpp — host-side planning of pairwise steps and whole contraction paths.
//
// What it replaces in the reference: the index bookkeeping Muscle.binary_einsum performs before it
// calls BLAS (classify inds into free/contracted/batch, permutedims + reshape), and the tree walk of
// Tangles.contract(tn; path) (call sites: /root/reference/src/Operations/overlap.jl:12,42-47,
// src/Algorithms/DMRG.jl:10-18).  Instead of materialising permuted operands, the planner emits
// additive offset tables (tnb_internal.h) and chooses the memory layout of every intermediate so that
// its single consumer reads it as a dense, M-fastest matrix.
//
// Pure host code: no CUDA calls here, so the same code backs the dry-run plans that CPU tests inspect.
#include <cstdlib>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include "tnb_internal.h"

namespace {

struct ModeUse {
    int64_t ext = 1;
    int64_t sa = 0, sb = 0, sc = 0;
    bool inA = false, inB = false, inC = false, inSum = false;
};

std::string fmt_mode(const char* what, int32_t mode) {
    char buf[128];
    snprintf(buf, sizeof buf, "%s (mode %d)", what, (int)mode);
    return buf;
}

// Build the two-level table of one index group for one tensor.
// modes are ordered fastest-first; strides[i] is the tensor's stride for modes[i] (0 if absent).
void build_table(const std::vector<int64_t>& ext, const std::vector<int64_t>& strides, HostTable* t,
                 int64_t lo_max = 4096) {
    const size_t n = ext.size();
    t->size = 1;
    for (size_t i = 0; i < n; i++) t->size *= ext[i];
    // lo prefix
    size_t npre = 0;
    int64_t lo = 1;
    const int64_t LO_MAX = lo_max;
    while (npre < n && (npre == 0 || lo * ext[npre] <= LO_MAX)) {
        lo *= ext[npre];
        npre++;
    }
    t->lo_size = lo;
    t->lo.assign((size_t)lo, 0);
    {
        std::vector<int64_t> dig(npre, 0);
        int64_t off = 0;
        for (int64_t i = 0; i < lo; i++) {
            t->lo[(size_t)i] = off;
            // increment mixed-radix counter
            for (size_t d = 0; d < npre; d++) {
                dig[d]++;
                off += strides[d];
                if (dig[d] < ext[d]) break;
                off -= strides[d] * ext[d];
                dig[d] = 0;
            }
        }
    }
    int64_t hi = t->size / lo;
    t->hi.assign((size_t)hi, 0);
    {
        const size_t nh = n - npre;
        std::vector<int64_t> dig(nh, 0);
        int64_t off = 0;
        for (int64_t i = 0; i < hi; i++) {
            t->hi[(size_t)i] = off;
            for (size_t d = 0; d < nh; d++) {
                dig[d]++;
                off += strides[npre + d];
                if (dig[d] < ext[npre + d]) break;
                off -= strides[npre + d] * ext[npre + d];
                dig[d] = 0;
            }
        }
    }
    // affine?
    t->affine = true;
    t->stride = 0;
    if (n > 0) {
        t->stride = strides[0];
        int64_t expect = strides[0];
        for (size_t d = 0; d < n; d++) {
            if (strides[d] != expect) { t->affine = false; break; }
            expect *= ext[d];
        }
        if (!t->affine) t->stride = 0;
    }
}

int64_t min_nonzero_stride(const std::vector<int64_t>& s) {
    int64_t best = INT64_MAX;
    for (int64_t v : s) {
        int64_t a = v < 0 ? -v : v;
        if (a != 0 && a < best) best = a;
    }
    return best;
}

}  // namespace

int tnb_build_step(const PlanTensor& A, const PlanTensor& B, const PlanTensor& C,
                   const std::vector<int32_t>& sum, bool cplx, size_t elem_size, StepSpec* out,
                   std::string* msg) {
    std::map<int32_t, ModeUse> use;
    auto add = [&](const PlanTensor& T, int which) -> int {
        std::set<int32_t> seen;
        for (size_t i = 0; i < T.modes.size(); i++) {
            int32_t m = T.modes[i];
            if (!seen.insert(m).second) {
                *msg = fmt_mode("repeated mode inside one tensor", m);
                return TNB_EINVAL;
            }
            if (T.ext[i] < 1) { *msg = fmt_mode("extent < 1", m); return TNB_EINVAL; }
            auto it = use.find(m);
            if (it == use.end()) {
                ModeUse u;
                u.ext = T.ext[i];
                it = use.emplace(m, u).first;
            } else if (it->second.ext != T.ext[i]) {
                *msg = fmt_mode("extent mismatch", m);
                return TNB_EINVAL;
            }
            ModeUse& u = it->second;
            if (which == 0) { u.inA = true; u.sa = T.stride[i]; }
            if (which == 1) { u.inB = true; u.sb = T.stride[i]; }
            if (which == 2) { u.inC = true; u.sc = T.stride[i]; }
        }
        return 0;
    };
    int rc;
    if ((rc = add(A, 0))) return rc;
    if ((rc = add(B, 1))) return rc;
    if ((rc = add(C, 2))) return rc;
    for (int32_t m : sum) {
        auto it = use.find(m);
        if (it == use.end()) continue;  // summing a mode nobody carries: no-op
        it->second.inSum = true;
    }
    struct Ent { int32_t mode; int64_t ext, sa, sb, sc; };
    std::vector<Ent> gm, gn, gk, gl;
    for (auto& kv : use) {
        const ModeUse& u = kv.second;
        Ent e{kv.first, u.ext, u.sa, u.sb, u.sc};
        if (!u.inA && !u.inB) { *msg = fmt_mode("output mode carried by neither operand", kv.first); return TNB_EINVAL; }
        if (u.inSum && u.inC) { *msg = fmt_mode("mode is both summed and in the output", kv.first); return TNB_EINVAL; }
        if (!u.inSum && !u.inC) { *msg = fmt_mode("mode is neither summed nor in the output", kv.first); return TNB_EINVAL; }
        if (u.ext == 1) continue;  // contributes nothing to addressing
        if (u.inSum) gk.push_back(e);
        else if (u.inA && u.inB) gl.push_back(e);
        else if (u.inA) gm.push_back(e);
        else gn.push_back(e);
    }
    auto by_c = [](const Ent& x, const Ent& y) {
        if (x.sc != y.sc) return x.sc < y.sc;
        return x.mode < y.mode;
    };
    // Free modes are enumerated in the order the OPERAND stores them (so that a dense operand reads as an
    // M-/N-fastest matrix and qualifies for the tensor-core kernels); the output is scattered through cm/cn.
    // For intermediates the planner lays the operand out in the output's order, so both orders coincide.
    std::stable_sort(gm.begin(), gm.end(), [](const Ent& x, const Ent& y) {
        if (x.sa != y.sa) return x.sa < y.sa;
        return x.mode < y.mode;
    });
    std::stable_sort(gn.begin(), gn.end(), [](const Ent& x, const Ent& y) {
        if (x.sb != y.sb) return x.sb < y.sb;
        return x.mode < y.mode;
    });
    std::stable_sort(gl.begin(), gl.end(), by_c);
    // K order: by A's stride where A carries the mode, B-only modes afterwards by B's stride.
    std::stable_sort(gk.begin(), gk.end(), [](const Ent& x, const Ent& y) {
        bool xa = x.sa != 0, ya = y.sa != 0;
        if (xa != ya) return xa;
        if (xa) { if (x.sa != y.sa) return x.sa < y.sa; }
        else if (x.sb != y.sb) return x.sb < y.sb;
        return x.mode < y.mode;
    });

    StepSpec& S = *out;
    auto mk = [&](const std::vector<Ent>& g, int which, HostTable* t) {
        std::vector<int64_t> ext, st;
        for (auto& e : g) {
            ext.push_back(e.ext);
            st.push_back(which == 0 ? e.sa : which == 1 ? e.sb : e.sc);
        }
        build_table(ext, st, t);
    };
    mk(gm, 0, &S.am); mk(gk, 0, &S.ak); mk(gl, 0, &S.al);
    mk(gn, 1, &S.bn); mk(gk, 1, &S.bk); mk(gl, 1, &S.bl);
    mk(gm, 2, &S.cm); mk(gn, 2, &S.cn); mk(gl, 2, &S.cl);
    S.M = S.am.size; S.N = S.bn.size; S.K = S.ak.size; S.L = S.al.size;
    if (S.M >= (1ll << 31) || S.N >= (1ll << 31) || S.K >= (1ll << 31) || S.L >= 65536) {
        *msg = "index group too large (M,N,K must be < 2^31, batch < 65536)";
        return TNB_EUNSUPPORTED;
    }
    S.conjA = A.conj; S.conjB = B.conj;
    {
        std::vector<int64_t> sm, sk;
        for (auto& e : gm) sm.push_back(e.sa);
        for (auto& e : gk) sk.push_back(e.sa);
        S.a_kfast = min_nonzero_stride(sk) < min_nonzero_stride(sm) ? 1 : 0;
        sm.clear(); sk.clear();
        for (auto& e : gn) sm.push_back(e.sb);
        for (auto& e : gk) sk.push_back(e.sb);
        S.b_kfast = min_nonzero_stride(sk) < min_nonzero_stride(sm) ? 1 : 0;
    }
    S.a_mmajor = S.L == 1 && S.am.affine && (S.M == 1 || S.am.stride == 1) && S.ak.affine &&
                 (S.K == 1 || S.ak.stride == S.M);
    S.b_nmajor = S.L == 1 && S.bn.affine && (S.N == 1 || S.bn.stride == 1) && S.bk.affine &&
                 (S.K == 1 || S.bk.stride == S.N);
    // "stem" steps: a huge dense operand times a tiny one (N <= 16, K <= 64) — HBM-bound.  The streaming kernel
    // walks the big free group in tiles of TM = lo_size of a re-grouped output table, so that every tile's
    // output addresses are  hi[tile] + (tile-invariant pattern);  the pattern is sorted once here (rel) together
    // with its inverse (pos), which lets the kernel write each tile in ascending address order (coalesced).
    S.st_ok = false;
    // Small operands whose width is a multiple of 128 (64 <= K <= 128) take 128 columns per pass (single accumulator
    // set, the big operand is read half as often, up to 16 passes); tried before the 64-column form (<= 8 passes).
    // TNB_STEM_N128=0 disables the 128-column form, TNB_STEM_MAX_PASSES overrides both pass limits (experiments).
    static const bool n128 = [] { const char* e = getenv("TNB_STEM_N128"); return e ? atoi(e) != 0 : true; }();
    for (int att = 0; att < 4 && !S.st_ok; att++) {
        const int sw = att >> 1;
        const bool wide = !(att & 1);
        if (wide && !n128) continue;
        const std::vector<Ent>& big = sw ? gn : gm;
        const std::vector<Ent>& small = sw ? gm : gn;
        const int64_t Mb = sw ? S.N : S.M, Ns = sw ? S.M : S.N;
        const bool dense = sw ? S.b_nmajor : S.a_mmajor;
        if (S.L != 1 || !dense || Mb < 65536) continue;
        // class 1: SIMT streaming (N*K small: FP32 FMA keeps up with HBM); class 2: tensor-core stem (c64 only)
        // the tensor-core stem kernel takes <= 64 small-side columns per pass; wider small operands (<= 256) run as
        // several passes that re-read the big operand (still far fewer bytes than a tile kernel without overlap)
        // K > 128 in the 128-column form (chunks of 128 k summed round-to-nearest): taken when the small operand is exactly
        // one pass wide (K <= 512) — the tile GEMM would write such an output 128 columns at a time, uncoalesced for every
        // layout but column-fastest (sycamore53_m14's 128 x 524288 x 512 step: 1.69 -> 1.12 ms).  With several passes the
        // tile GEMM re-reads less and wins; TNB_STEM_KMAX=512 sends those here as well (experiment, measured slower).
        static const int64_t kmax_env = [] { const char* e = getenv("TNB_STEM_KMAX"); return e ? atoll(e) : 128ll; }();
        const int64_t kmax_wide = Ns == 128 ? std::max<int64_t>(kmax_env, 512) : kmax_env;
        if (wide && (Ns % 128 != 0 || S.K < 64 || S.K > kmax_wide)) continue;
        const int64_t nper = wide ? 128 : (Ns > 64 ? 64 : Ns);
        const int64_t npass = Ns / std::max<int64_t>(nper, 1);
        static const int64_t max_pass_env = [] { const char* e = getenv("TNB_STEM_MAX_PASSES"); return e ? atoll(e) : 0ll; }();
        const int64_t max_pass = max_pass_env > 0 ? max_pass_env : (wide ? 16 : 8);
        const bool tcst = cplx && elem_size == 8 && Ns % nper == 0 && npass <= max_pass && tnb_stem_tc_shape_ok(Mb, nper, S.K);
        const bool simt = !tcst && Ns <= 16 && S.K <= 64;
        if (!simt && !tcst) continue;
        const int64_t lo_max = simt ? std::max<int64_t>(64, 2048 / std::max<int64_t>(Ns, 1)) : 128;
        std::vector<int64_t> ext, st;
        for (auto& e : big) {
            // a single mode longer than a tile is split (d, ext/d) with d the largest divisor that fits
            if (ext.empty() && e.ext > lo_max) {
                int64_t d = 1;
                for (int64_t c = 2; c <= lo_max; c++) if (e.ext % c == 0) d = c;
                ext.push_back(d); st.push_back(e.sc);
                ext.push_back(e.ext / d); st.push_back(e.sc * d);
                continue;
            }
            ext.push_back(e.ext); st.push_back(e.sc);
        }
        HostTable cb;
        build_table(ext, st, &cb, lo_max);
        const int64_t TM = cb.lo_size;
        if (TM < 64 || TM > 4096 || TM * (simt ? Ns : nper) > (wide ? 16384 : 8192) || (TM % 2)) continue;
        if (!simt && TM != 128) continue;
        S.st_tc = !simt;
        std::vector<int64_t> ext2, st2;
        for (auto& e : small) { ext2.push_back(e.ext); st2.push_back(e.sc); }
        HostTable cs;
        build_table(ext2, st2, &cs);
        const int64_t ncol = simt ? Ns : nper, passes = simt ? 1 : npass;
        const int64_t cnt = TM * ncol;
        S.st_rel.assign((size_t)(cnt * passes), 0);
        S.st_pos.assign((size_t)(cnt * passes), 0);
        bool contig = true;
        int64_t R = cnt;
        for (int64_t ps = 0; ps < passes; ps++) {
            std::vector<std::pair<int64_t, int64_t>> addr((size_t)cnt);
            for (int64_t ml = 0; ml < TM; ml++)
                for (int64_t n = 0; n < ncol; n++)
                    addr[(size_t)(ml * ncol + n)] = {cb.lo[(size_t)ml] + cs.at(ps * ncol + n), ml * ncol + n};
            std::sort(addr.begin(), addr.end());
            int64_t* rel = S.st_rel.data() + ps * cnt;
            int64_t* pos = S.st_pos.data() + ps * cnt;
            for (int64_t j = 0; j < cnt; j++) {
                rel[j] = addr[(size_t)j].first;
                pos[addr[(size_t)j].second] = j;
                if (addr[(size_t)j].first != addr[0].first + j) contig = false;
            }
            // contiguous runs: largest power-of-two r with rel[j] == rel[j - j%r] + j%r for all j
            int64_t r = 1;
            while (r * 2 <= cnt && cnt % (r * 2) == 0) {
                const int64_t r2 = r * 2;
                bool ok = true;
                for (int64_t j = 0; j < cnt && ok; j += r2)
                    for (int64_t o = r; o < r2; o++)
                        if (rel[j + o] != rel[j] + o) { ok = false; break; }
                if (!ok) break;
                r = r2;
            }
            R = std::min(R, r);
        }
        S.st_run = (int32_t)R;
        // rank separable in (row, column) on disjoint bits?  (always true for power-of-two extents: the rank is a bit deposit)
        bool additive = getenv("TNB_STEM_NO_ADDITIVE") == nullptr, even = R >= 2;   // env: test hook for the table path
        for (int64_t ps = 0; ps < passes && additive; ps++) {
            const int64_t* pos = S.st_pos.data() + ps * cnt;
            for (int64_t ml = 0; ml < TM && additive; ml++)
                for (int64_t n = 0; n < ncol; n++)
                    if (pos[ml * ncol + n] != pos[ml * ncol] + pos[n] - pos[0] || (pos[ml * ncol] & (pos[n] - pos[0])) != 0) { additive = false; break; }
        }
        for (int64_t v : cb.hi) if (v & 1) even = false;
        for (int64_t j = 0; j < cnt * passes && even; j += R) if (S.st_rel[(size_t)j] & 1) even = false;
        if (wide && !additive) continue;       // the 128-column form has no room for the general rank table
        S.st_additive = additive; S.st_even = even;
        S.st_rel_small = true;
        for (int64_t v : S.st_rel) if (v < 0 || v >= ((int64_t)1 << 31)) { S.st_rel_small = false; break; }
        // Direct epilogue of the stem kernels (tensor-core forms: <= 64 columns per pass on the 1-CTA kernel, 128-column passes
        // on the CTA-pair kernel; SIMT form: N <= 16): a thread owns one row of the tile and a warp stores one column of 32
        // consecutive rows per instruction.  That is as good as the staged, sorted write-out whenever those 32 addresses are
        // whole 32-byte sectors (4 complex64 / 2 complex128) in <= 4 lines per instruction: the big side's fastest rows are
        // the output's fastest index — then the shared-memory round trip of the staging tile is skipped altogether.
        // TNB_STEM_DIRECT (read at plan time: test hook): 0 keeps every step on the staged path, 1 asks for whole 64-byte
        // pieces (the first version; 128 x 8388608 x 128 of the committed path gains 3 % from the 32-byte rule: its rows
        // fill sectors 0-1 / 2-3 of a line in alternate columns); TNB_STEM_NO_ADDITIVE (the test hook of the general rank
        // table) implies 0.  st_pairs: rows 2i, 2i+1 are adjacent in the output (the SIMT complex64 form stores 16 bytes).
        S.st_direct = false; S.st_pairs = false;
        {
            const char* e = getenv("TNB_STEM_DIRECT");
            const int mode = e ? atoi(e) : 2;
            const int64_t pe = (mode == 1 ? 64 : 32) / (int64_t)elem_size;       // elements per piece (modes 2, 3: one sector)
            bool ok = S.st_rel_small && TM % 32 == 0 && mode != 0 && pe >= 1 && getenv("TNB_STEM_NO_ADDITIVE") == nullptr;
            bool pairs = TM % 2 == 0;
            const int64_t line = 128 / (int64_t)elem_size;
            for (int64_t v : cb.hi) { if (v % pe) ok = false; if (v & 1) pairs = false; }
            for (int64_t ps = 0; ps < passes && ok; ps++) {
                const int64_t* rel = S.st_rel.data() + ps * cnt;
                const int64_t* pos = S.st_pos.data() + ps * cnt;
                for (int64_t n = 0; n < ncol && ok; n++) {
                    if (!(simt && mode >= 2) && (rel[pos[n]] - rel[pos[0]]) % pe) ok = false;
                    if ((rel[pos[n]] - rel[pos[0]]) & 1) pairs = false;
                }
                // SIMT form: the rule is applied to ALL columns of the 32 rows together — a thread stores its N results back
                // to back, so a small-operand index below the rows (the MPO bond of configs[4]) still fills whole sectors within
                // a few instructions of the same warp (1024^2 x 6 x 6 complex128: 2.77 -> 3.12 TB/s; TNB_STEM_DIRECT=1 keeps
                // the per-column rule for this form too)
                const bool all_cols = simt && mode >= 2;
                for (int64_t q = 0; q < TM / 32 && ok; q++) {
                    std::vector<int64_t> piece;
                    std::set<int64_t> lines;
                    for (int64_t l = 0; l < 32; l++) {
                        for (int64_t n = 0; n < (all_cols ? ncol : 1); n++) {
                            const int64_t a = rel[pos[(q * 32 + l) * ncol + n]];
                            piece.push_back(a / pe);
                            if (n == 0) lines.insert(a / line);
                        }
                    }
                    std::sort(piece.begin(), piece.end());
                    for (size_t i = 0; i < piece.size() && ok; i += (size_t)pe)
                        if (piece[i] != piece[i + (size_t)pe - 1] || (i + (size_t)pe < piece.size() && piece[i + (size_t)pe] == piece[i])) ok = false;
                    if (lines.size() > (all_cols ? 16u : 4u)) ok = false;
                }
                for (int64_t ml = 0; ml + 1 < TM && pairs; ml += 2) {
                    const int64_t a0 = rel[pos[ml * ncol]], a1 = rel[pos[(ml + 1) * ncol]];
                    if (a1 != a0 + 1 || (a0 & 1)) pairs = false;
                }
            }
            S.st_direct = ok; S.st_pairs = ok && pairs;
            if (getenv("TNB_DEBUG_STEM") && !ok)
                fprintf(stderr, "[stem-direct] no: rel_small=%d TM=%lld pe=%lld hi0=%lld hi1=%lld col1=%lld row1=%lld row32=%lld\n", (int)S.st_rel_small, (long long)TM, (long long)pe,
                        (long long)cb.hi[0], (long long)(cb.hi.size() > 1 ? cb.hi[1] : 0), (long long)(ncol > 1 ? S.st_rel[S.st_pos[1]] - S.st_rel[S.st_pos[0]] : 0),
                        (long long)(S.st_rel[S.st_pos[ncol]] - S.st_rel[S.st_pos[0]]), (long long)(TM > 32 ? S.st_rel[S.st_pos[32 * ncol]] - S.st_rel[S.st_pos[0]] : 0));
        }
        if (getenv("TNB_DEBUG_STEM")) {
            // which bits of the in-tile rank belong to the big-side rows / the small-side columns (pass 0)
            int64_t rowbits = 0, colbits = 0;
            for (int64_t ml = 0; ml < TM; ml++) rowbits |= S.st_pos[(size_t)(ml * ncol)] - S.st_pos[0];
            for (int64_t n = 0; n < ncol; n++) colbits |= S.st_pos[(size_t)n] - S.st_pos[0];
            fprintf(stderr, "[stem] M=%lld N=%lld K=%lld big=%lld small=%lld tc=%d swap=%d TM=%lld ncol=%lld passes=%lld run=%lld additive=%d even=%d contig=%d direct=%d rowbits=0x%llx colbits=0x%llx\n",
                    (long long)S.M, (long long)S.N, (long long)S.K, (long long)Mb, (long long)Ns, (int)!simt, sw, (long long)TM,
                    (long long)ncol, (long long)passes, (long long)R, (int)additive, (int)even, (int)contig, (int)S.st_direct,
                    (unsigned long long)rowbits, (unsigned long long)colbits);
        }
        S.st_npass = (int32_t)passes; S.st_ncol = (int32_t)ncol;
        S.st_hi = cb.hi;
        S.st_ok = true; S.st_swap = sw != 0; S.st_tm = (int32_t)TM; S.st_contig = contig;
    }
    double macs = (double)S.M * (double)S.N * (double)S.K * (double)S.L;
    S.flops = (cplx ? 8.0 : 2.0) * macs;
    auto elems = [](const PlanTensor& T) {
        int64_t e = 1;
        for (int64_t x : T.ext) e *= x;
        return e;
    };
    S.a_elems = elems(A); S.b_elems = elems(B); S.c_elems = elems(C);
    S.bytes = (double)elem_size * ((double)S.a_elems + (double)S.b_elems + (double)S.c_elems);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Whole-path planning
// ------------------------------------------------------------------------------------------------
int tnb_plan_build(const tnb_tensor* leaves, int32_t nleaves, const int32_t* steps, int32_t nsteps,
                   const int32_t* sliced_modes, int32_t nsliced, const tnb_tensor* out, tnb_plan* plan,
                   std::string* msg) {
    if (nleaves < 1 || nsteps != nleaves - 1) {
        *msg = "a path over n leaves needs exactly n-1 pairwise steps (join disconnected parts with outer-product steps)";
        return TNB_EINVAL;
    }
    if (!out) { *msg = "out descriptor is NULL"; return TNB_EINVAL; }
    const int dtype = out->dtype;
    const size_t esz = tnb_dtype_size(dtype);
    if (!esz) { *msg = "unsupported dtype"; return TNB_EUNSUPPORTED; }
    const bool cplx = tnb_dtype_complex(dtype);
    plan->dtype = dtype;
    plan->nleaves = nleaves;
    plan->nsteps = nsteps;
    const int nn = nleaves + nsteps;
    plan->nodes.assign(nn, PlanNode());
    plan->steps.assign(nsteps, StepSpec());

    std::set<int32_t> sliced(sliced_modes, sliced_modes + nsliced);
    if ((int)sliced.size() != nsliced) { *msg = "repeated sliced mode"; return TNB_EINVAL; }
    plan->sliced_modes.assign(sliced_modes, sliced_modes + nsliced);
    plan->sliced_ext.assign(nsliced, 0);

    std::map<int32_t, int64_t> ext_of;
    std::map<int32_t, int> total;
    plan->leaf_buf.resize(nleaves);
    plan->leaf_off.resize(nleaves);
    for (int i = 0; i < nleaves; i++) {
        const tnb_tensor& T = leaves[i];
        if (T.dtype != dtype) { *msg = "all leaves and the output must share one dtype (promote on the host)"; return TNB_EINVAL; }
        if (T.rank < 0 || T.rank > TNB_MAX_RANK) { *msg = "leaf rank out of range"; return TNB_EUNSUPPORTED; }
        PlanNode& nd = plan->nodes[i];
        nd.leaf = true;
        nd.t.conj = T.conj;
        nd.t.offset = T.offset_elems;
        nd.slice_stride.assign(nsliced, 0);
        if (nsliced == 0) nd.dep = true;   // un-sliced network = one slice; nothing is "hoisted"
        plan->leaf_buf[i] = T.buf;
        plan->leaf_off[i] = T.offset_elems;
        std::set<int32_t> seen;
        for (int r = 0; r < T.rank; r++) {
            int32_t m = T.mode[r];
            if (!seen.insert(m).second) { *msg = fmt_mode("repeated mode inside a leaf", m); return TNB_EINVAL; }
            auto it = ext_of.find(m);
            if (it == ext_of.end()) ext_of[m] = T.extent[r];
            else if (it->second != T.extent[r]) { *msg = fmt_mode("extent mismatch between leaves", m); return TNB_EINVAL; }
            if (T.extent[r] < 1) { *msg = fmt_mode("extent < 1", m); return TNB_EINVAL; }
            if (sliced.count(m)) {
                for (int j = 0; j < nsliced; j++)
                    if (sliced_modes[j] == m) {
                        nd.slice_stride[j] = T.stride_elems[r];
                        plan->sliced_ext[j] = T.extent[r];
                    }
                nd.dep = true;
                continue;
            }
            nd.t.modes.push_back(m);
            nd.t.ext.push_back(T.extent[r]);
            nd.t.stride.push_back(T.stride_elems[r]);
            nd.elems *= T.extent[r];
            total[m]++;
        }
    }
    plan->nslices = 1;
    for (int j = 0; j < nsliced; j++) {
        if (plan->sliced_ext[j] == 0) { *msg = fmt_mode("sliced mode not carried by any leaf", sliced_modes[j]); return TNB_EINVAL; }
        plan->nslices *= plan->sliced_ext[j];
    }
    // output
    if (out->rank < 0 || out->rank > TNB_MAX_RANK) { *msg = "output rank out of range"; return TNB_EUNSUPPORTED; }
    PlanTensor outT;
    for (int r = 0; r < out->rank; r++) {
        int32_t m = out->mode[r];
        if (sliced.count(m)) { *msg = fmt_mode("an output mode cannot be sliced", m); return TNB_EINVAL; }
        auto it = ext_of.find(m);
        if (it == ext_of.end()) { *msg = fmt_mode("output mode not carried by any leaf", m); return TNB_EINVAL; }
        if (it->second != out->extent[r]) { *msg = fmt_mode("output extent mismatch", m); return TNB_EINVAL; }
        outT.modes.push_back(m);
        outT.ext.push_back(out->extent[r]);
        outT.stride.push_back(out->stride_elems[r]);
        total[m]++;
    }
    plan->out_buf = out->buf;
    plan->out_off = out->offset_elems;

    // bottom-up: mode sets of intermediates
    std::vector<std::map<int32_t, int>> cnt(nn);
    for (int i = 0; i < nleaves; i++)
        for (int32_t m : plan->nodes[i].t.modes) cnt[i][m] = 1;
    std::vector<std::vector<int32_t>> summed(nsteps);
    for (int s = 0; s < nsteps; s++) {
        int a = steps[2 * s], b = steps[2 * s + 1], c = nleaves + s;
        if (a < 0 || b < 0 || a >= c || b >= c || a == b) { *msg = "step references an id that does not exist yet"; return TNB_EINVAL; }
        if (plan->nodes[a].consumer >= 0 || plan->nodes[b].consumer >= 0) { *msg = "an id is consumed by two steps"; return TNB_EINVAL; }
        plan->nodes[a].consumer = c;
        plan->nodes[b].consumer = c;
        PlanNode& nd = plan->nodes[c];
        nd.a = a; nd.b = b;
        nd.dep = plan->nodes[a].dep || plan->nodes[b].dep;
        cnt[c] = cnt[a];
        for (auto& kv : cnt[b]) cnt[c][kv.first] += kv.second;
        for (auto& kv : cnt[c]) {
            if (kv.second < total[kv.first]) {
                nd.t.modes.push_back(kv.first);
                nd.t.ext.push_back(ext_of[kv.first]);
                nd.elems *= ext_of[kv.first];
            } else {
                summed[s].push_back(kv.first);
            }
        }
        cnt[a].clear();
        cnt[b].clear();
        plan->steps[s].a_id = a; plan->steps[s].b_id = b; plan->steps[s].c_id = c;
        plan->steps[s].hoisted = !nd.dep;
    }
    const int root = nn - 1;
    if (nsteps > 0) {
        std::set<int32_t> rm(plan->nodes[root].t.modes.begin(), plan->nodes[root].t.modes.end());
        std::set<int32_t> om(outT.modes.begin(), outT.modes.end());
        if (rm != om) { *msg = "output modes differ from the open modes of the path result"; return TNB_EINVAL; }
        plan->nodes[root].t = outT;
    } else {
        *msg = "single-leaf network: nothing to contract";
        return TNB_EINVAL;
    }
    if (nsliced > 0 && !plan->nodes[root].dep) { *msg = "sliced modes do not reach the root"; return TNB_EINVAL; }

    // top-down: choose intermediate layouts. For child X of step (A,B)->C:
    //   X = [ free modes of X in C's order (fastest) | summed modes in the step's K order | batch modes in C's order ]
    for (int s = nsteps - 1; s >= 0; s--) {
        const int a = plan->steps[s].a_id, b = plan->steps[s].b_id, c = plan->steps[s].c_id;
        const PlanTensor& Ct = plan->nodes[c].t;
        std::vector<size_t> corder(Ct.modes.size());
        for (size_t i = 0; i < corder.size(); i++) corder[i] = i;
        std::stable_sort(corder.begin(), corder.end(), [&](size_t x, size_t y) {
            if (Ct.stride[x] != Ct.stride[y]) return Ct.stride[x] < Ct.stride[y];
            return Ct.modes[x] < Ct.modes[y];
        });
        std::set<int32_t> inA(plan->nodes[a].t.modes.begin(), plan->nodes[a].t.modes.end());
        std::set<int32_t> inB(plan->nodes[b].t.modes.begin(), plan->nodes[b].t.modes.end());
        // K order
        std::vector<int32_t> korder;
        {
            const PlanNode* ref = plan->nodes[a].leaf ? &plan->nodes[a] : plan->nodes[b].leaf ? &plan->nodes[b] : nullptr;
            std::set<int32_t> ks(summed[s].begin(), summed[s].end());
            if (ref) {
                std::vector<std::pair<int64_t, int32_t>> v;
                for (size_t i = 0; i < ref->t.modes.size(); i++)
                    if (ks.count(ref->t.modes[i])) v.push_back({ref->t.stride[i], ref->t.modes[i]});
                std::sort(v.begin(), v.end());
                for (auto& p : v) { korder.push_back(p.second); ks.erase(p.second); }
            }
            for (int32_t m : ks) korder.push_back(m);
        }
        for (int side = 0; side < 2; side++) {
            const int x = side == 0 ? a : b;
            PlanNode& X = plan->nodes[x];
            if (X.leaf) continue;
            const std::set<int32_t>& mine = side == 0 ? inA : inB;
            const std::set<int32_t>& other = side == 0 ? inB : inA;
            std::vector<int32_t> lay;
            for (size_t i : corder) {
                int32_t m = Ct.modes[i];
                if (mine.count(m) && !other.count(m)) lay.push_back(m);
            }
            for (int32_t m : korder) if (mine.count(m)) lay.push_back(m);
            for (size_t i : corder) {
                int32_t m = Ct.modes[i];
                if (mine.count(m) && other.count(m)) lay.push_back(m);
            }
            if (lay.size() != X.t.modes.size()) { *msg = "internal: layout size mismatch"; return TNB_EINVAL; }
            X.t.modes = lay;
            X.t.ext.clear(); X.t.stride.clear();
            int64_t st = 1;
            for (int32_t m : lay) {
                X.t.ext.push_back(ext_of[m]);
                X.t.stride.push_back(st);
                st *= ext_of[m];
            }
        }
    }
    // per-step GEMM views
    plan->info = tnb_plan_info{};
    plan->info.nslices = plan->nslices;
    for (int s = 0; s < nsteps; s++) {
        StepSpec& S = plan->steps[s];
        int rc = tnb_build_step(plan->nodes[S.a_id].t, plan->nodes[S.b_id].t, plan->nodes[S.c_id].t, summed[s],
                                cplx, esz, &S, msg);
        if (rc) return rc;
        if (S.hoisted) {
            plan->order_hoisted.push_back(s);
            plan->info.nsteps_hoisted++;
            plan->info.flops_hoisted += S.flops;
            plan->info.bytes_hoisted += S.bytes;
        } else {
            plan->order_dep.push_back(s);
            plan->info.nsteps_per_slice++;
            plan->info.flops_per_slice += S.flops;
            plan->info.bytes_per_slice += S.bytes;
        }
        if (S.c_id != root) plan->info.max_intermediate_elems = std::max(plan->info.max_intermediate_elems, S.c_elems);
    }

    // arena: first-fit interval allocation over the execution timeline (hoisted steps, then dep steps)
    {
        const int64_t align = std::max<int64_t>(1, 1024 / (int64_t)esz);
        std::vector<int> when(nn, -1);
        int t = 0;
        for (int s : plan->order_hoisted) when[nleaves + s] = t++;
        for (int s : plan->order_dep) when[nleaves + s] = t++;
        const int T_END = t + 1;
        struct Live { int64_t off, size; int death; };
        std::vector<Live> live;
        std::vector<int> by_birth;
        for (int s : plan->order_hoisted) by_birth.push_back(nleaves + s);
        for (int s : plan->order_dep) by_birth.push_back(nleaves + s);
        int64_t top = 0;
        for (int id : by_birth) {
            if (id == root) continue;
            PlanNode& nd = plan->nodes[id];
            int birth = when[id];
            int death = when[nd.consumer];
            if (!nd.dep && plan->nodes[nd.consumer].dep) death = T_END;  // hoisted value reused by every slice
            live.erase(std::remove_if(live.begin(), live.end(), [&](const Live& l) { return l.death < birth; }), live.end());
            std::sort(live.begin(), live.end(), [](const Live& x, const Live& y) { return x.off < y.off; });
            int64_t size = (nd.elems + align - 1) / align * align;
            int64_t pos = 0;
            for (const Live& l : live) {
                if (pos + size <= l.off) break;
                pos = std::max(pos, l.off + l.size);
            }
            nd.arena_off = pos;
            live.push_back({pos, size, death});
            top = std::max(top, pos + size);
        }
        plan->arena_elems = top;
        plan->info.workspace_bytes = top * (int64_t)esz;
    }
    return 0;
}
