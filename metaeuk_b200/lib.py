"""ctypes binding of libmetaeuk_b200.so (C ABI: include/metaeuk_b200.h) and the build recipe for it."""
import ctypes as C
import os
import subprocess

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB_PATH = os.path.join(PKG, "libmetaeuk_b200.so")
CSRC = os.path.join(PKG, "csrc")

HIT_DTYPE = np.dtype([("seq_id", "<u4"), ("score", "<i4"), ("diag", "<i4")])
ALN_DTYPE = np.dtype([("seq_id", "<u4"), ("raw_score", "<i4"), ("bit_score", "<i4"), ("q_start", "<i4"), ("q_end", "<i4"), ("db_start", "<i4"),
                      ("db_end", "<i4"), ("aln_len", "<i4"), ("passed", "<i4"), ("seq_id_est", "<f4"), ("q_cov", "<f4"), ("db_cov", "<f4"),
                      ("evalue", "<f8")], align=True)
CAND_DTYPE = np.dtype([("target_key", "<u4"), ("bit_score", "<i4"), ("q_start", "<i4"), ("q_end", "<i4"), ("t_start", "<i4"), ("t_end", "<i4"),
                       ("t_len", "<i4"), ("orf_key", "<u4"), ("frag_from", "<i4"), ("frag_to", "<i4")])
EXON_DTYPE = np.dtype([("cand_idx", "<u4"), ("contig_start", "<i4"), ("contig_end", "<i4"), ("nuc_len", "<i4")])
CALL_DTYPE = np.dtype([("target_key", "<u4"), ("strand", "<i4"), ("total_bit_score", "<i4"), ("n_exons", "<i4"), ("exon_begin", "<u4"),
                       ("low_coord", "<i4"), ("high_coord", "<i4"), ("pad", "<i4"), ("evalue", "<f8")], align=True)


class Params(C.Structure):
    _fields_ = [("kmer_thr", C.c_int32), ("max_res_list_len", C.c_int32), ("min_diag_score", C.c_int32), ("bin_size", C.c_int32),
                ("gap_open", C.c_int32), ("gap_extend", C.c_int32), ("min_aln_len", C.c_int32), ("pad0", C.c_int32), ("eval_thr", C.c_double),
                ("min_intron", C.c_int32), ("max_intron", C.c_int32), ("min_exon_aa", C.c_int32), ("max_overlap", C.c_int32),
                ("set_gap_open", C.c_int32), ("set_gap_extend", C.c_int32), ("max_exon_sets", C.c_int32), ("pad1", C.c_int32),
                ("tcov", C.c_double), ("metaeuk_eval", C.c_double),
                ("cap_items_log2", C.c_int32), ("cap_kmers_log2", C.c_int32), ("cap_hits_log2", C.c_int32), ("pad2", C.c_int32),
                ("comp_bias_corr", C.c_int32), ("reverse_fragments", C.c_int32), ("translation_table", C.c_int32), ("pad3", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("kmers", "lookups", "db_matches", "double_hits", "diag_cells", "pref_results", "sw_pairs", "sw_cells_fwd",
                                           "sw_cells_rev", "sw_target_bytes", "query_residues", "n_queries")] + \
               [(n, C.c_double) for n in ("ms_prep", "ms_prefilter", "ms_sw_fwd", "ms_sw_rev", "ms_chain", "ms_h2d", "ms_d2h", "ms_host")] + \
               [("launches", C.c_uint32), ("overflow_queries", C.c_uint32), ("pref_phase_cycles", C.c_uint64 * 16), ("ms_orf", C.c_double), ("ms_orf_host", C.c_double),
                ("orf_nucleotides", C.c_uint64), ("orf_fragments", C.c_uint64), ("orf_residues", C.c_uint64), ("ms_reduce", C.c_double),
                ("ms_join", C.c_double), ("ms_predict", C.c_double)] + \
               [(n, C.c_uint64) for n in ("pe_batches", "pe_contigs", "pe_fragments", "pe_alignments", "pe_calls", "pe_exons", "pe_streams")] + \
               [(n, C.c_double) for n in ("ms_prefilter_sum", "ms_sw_fwd_sum", "ms_sw_rev_sum", "ms_dev_sum", "ms_dev_busy")]
PEXON_DTYPE = np.dtype([("orf_key", "<u4"), ("bit_score", "<i4"), ("seq_id_est", "<f4"), ("t_start", "<i4"), ("t_end", "<i4"), ("t_len", "<i4"),
                        ("contig_start", "<i4"), ("contig_end", "<i4"), ("nuc_len", "<i4"), ("frag_from", "<i4"), ("frag_to", "<i4"), ("pad", "<i4"), ("evalue", "<f8")])
PRED_DTYPE = np.dtype([("target_key", "<u4"), ("strand", "<i4"), ("total_bit_score", "<u4"), ("n_exons", "<u4"), ("low_coord", "<u4"), ("high_coord", "<u4"),
                       ("exon_begin", "<u4"), ("pad", "<u4"), ("evalue", "<f8")])
PRED_OUT_DTYPE = np.dtype([("cluster_id", "<u4"), ("cluster_low_coord", "<u4"), ("no_overlap_cluster_id", "<u4"), ("is_rep", "<u4")])
ORF_DTYPE = np.dtype([("contig", "<u4"), ("from", "<u4"), ("to", "<u4"), ("flags", "<u4"), ("aa_off", "<u8"), ("aa_len", "<u4"), ("strand_from", "<u4")])


class Mb200Error(RuntimeError):
    pass


def _gxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build_library(force=False, verbose=False):
    """nvcc cross-compiles for sm_100a without a GPU; the .so stays in-tree (it travels with the repo snapshot)."""
    srcs = [os.path.join(CSRC, f) for f in ("mb200.cu", "host_tables.cpp", "host_join.cpp", "host_indexdb.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("prefilter_core.cuh", "chain_core.cuh", "reduce_core.cuh", "orf_kernels.cuh", "sw_kernels.cuh", "sw_stream.cuh", "cta_prims.cuh", "host_tables.h", "host_join.h", "host_indexdb.h", "submat_data.h")] + \
        [os.path.join(ROOT, "include", "metaeuk_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ["nvcc", "-ccbin", _gxx(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off", "-shared"] + srcs + ["-o", LIB_PATH, "-lgomp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def load_library():
    """Load the in-tree library; fail loudly if it is missing (there is no CPU path)."""
    global _lib
    if _lib is None:
        path = os.environ.get("MB200_LIB_PATH", LIB_PATH)   # (tuning experiments load an alternative build of the same sources)
        if not os.path.exists(path):
            raise Mb200Error("libmetaeuk_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                             "metaeuk_b200 has no CPU fallback")
        L = C.CDLL(path)
        L.mb200_last_error.restype = C.c_char_p
        L.mb200_evalue.restype = C.c_double
        L.mb200_evalue.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.mb200_bit_score.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def build_kind():
    """"sm_100a" for the library; "simt-emulation" only when a test pointed MB200_LIB_PATH at tests/simt (kernels as CPU fibers)"""
    L = load_library()
    L.mb200_build_kind.restype = C.c_char_p
    return L.mb200_build_kind().decode()


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def pack_codes(seqs):
    """list of uint8 code arrays -> (concatenated, uint64 offsets)"""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    data = np.concatenate(seqs).astype(np.uint8, copy=False) if len(seqs) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(data), off


def mask_targets(unmasked, off, mask_prob=0.9, simd_width=4):
    """tantan masking of the targets as the reference's index builder does it (host only): -> (masked codes, number of masked residues)"""
    L = load_library()
    unmasked = np.ascontiguousarray(unmasked, np.uint8); off = np.ascontiguousarray(off, np.uint64)
    out = np.zeros(max(1, len(unmasked)), np.uint8); n = C.c_uint64(0)
    rc = L.mb200_mask_targets(_p(unmasked if len(unmasked) else np.zeros(1, np.uint8)), _p(off), C.c_uint32(len(off) - 1), C.c_double(mask_prob), int(simd_width), _p(out), C.byref(n))
    if rc != 0:
        raise Mb200Error("mb200 error %d: %s" % (rc, L.mb200_last_error().decode()))
    return out[:len(unmasked)], int(n.value)


def mask_targets_ex(unmasked, off, mask_tantan=True, mask_prob=0.9, lower_case=None, mask_n_repeat=0, simd_width=4):
    """Masker::maskSequence with all of its switches (--mask, --mask-prob, --mask-lower-case via one byte per residue, --mask-n-repeat): host only
    -> (masked codes, the reference's "Masked residues" count)"""
    L = load_library()
    unmasked = np.ascontiguousarray(unmasked, np.uint8); off = np.ascontiguousarray(off, np.uint64)
    lc = np.ascontiguousarray(lower_case, np.uint8) if lower_case is not None else None
    out = np.zeros(max(1, len(unmasked)), np.uint8); n = C.c_uint64(0)
    rc = L.mb200_mask_targets_ex(_p(unmasked if len(unmasked) else np.zeros(1, np.uint8)), _p(off), C.c_uint32(len(off) - 1), int(bool(mask_tantan)), C.c_double(mask_prob),
                                 int(simd_width), _p(lc) if lc is not None else None, int(mask_n_repeat), _p(out), C.byref(n))
    if rc != 0:
        raise Mb200Error("mb200 error %d: %s" % (rc, L.mb200_last_error().decode()))
    return out[:len(unmasked)], int(n.value)


def index_file_write(path, masked, unmasked, off, keys=None, kmer_thr=127, mask_mode=1):
    """build the target index on the host and persist it in the reference's index-DB layout (no GPU needed)"""
    L = load_library()
    masked = np.ascontiguousarray(masked, np.uint8); unmasked = np.ascontiguousarray(unmasked, np.uint8); off = np.ascontiguousarray(off, np.uint64)
    k = np.ascontiguousarray(keys, np.uint32) if keys is not None else None
    rc = L.mb200_index_file_write(os.fsencode(path), _p(masked), _p(unmasked), _p(off), C.c_uint32(len(off) - 1), _p(k) if k is not None else None, int(kmer_thr), int(mask_mode))
    if rc != 0:
        raise Mb200Error("mb200 error %d: %s" % (rc, L.mb200_last_error().decode()))


class Context:
    """One GPU context (mb200_ctx).  All inputs are numeric residue codes (see metaeuk_b200.modules.encode)."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        self._check(self.L.mb200_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.params = Params()
        self.L.mb200_params_default(C.byref(self.params))
        self.n_targets = 0

    def _check(self, rc):
        if rc != 0:
            raise Mb200Error("mb200 error %d: %s" % (rc, self.L.mb200_last_error().decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.mb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise KeyError(k)
            setattr(self.params, k, v)
        self._check(self.L.mb200_params_set(self.h, C.byref(self.params)))

    def matrices_default(self):
        self._check(self.L.mb200_matrices_default(self.h))

    def matrices_get(self):
        s = np.zeros(441, np.int16); u = np.zeros(441, np.int16); a = np.zeros(441, np.int16)
        self._check(self.L.mb200_matrices_get(self.h, _p(s), _p(u), _p(a)))
        return s, u, a

    def index_build(self, masked, unmasked, off):
        """masked/unmasked: concatenated uint8 codes, off: uint64[n+1]"""
        masked = np.ascontiguousarray(masked, np.uint8); unmasked = np.ascontiguousarray(unmasked, np.uint8); off = np.ascontiguousarray(off, np.uint64)
        self.n_targets = len(off) - 1
        self.target_off = off
        self._check(self.L.mb200_index_build(self.h, _p(masked), _p(unmasked), _p(off), C.c_uint32(self.n_targets)))

    def index_import(self, offsets_u64, entries6, masked, unmasked, off):
        offsets_u64 = np.ascontiguousarray(offsets_u64, np.uint64); entries6 = np.ascontiguousarray(entries6, np.uint8)
        masked = np.ascontiguousarray(masked, np.uint8); unmasked = np.ascontiguousarray(unmasked, np.uint8); off = np.ascontiguousarray(off, np.uint64)
        self.n_targets = len(off) - 1
        self.target_off = off
        self._check(self.L.mb200_index_import(self.h, _p(offsets_u64), _p(entries6), C.c_uint64(len(entries6) // 6), _p(masked), _p(unmasked), _p(off),
                                              C.c_uint32(self.n_targets)))

    # ---- persisted index (createindex layout) ----
    def index_save(self, path):
        self._check(self.L.mb200_index_save(self.h, os.fsencode(path)))

    def index_load(self, path):
        self._check(self.L.mb200_index_load(self.h, os.fsencode(path)))
        self.n_targets = self.index_sizes()[1]

    def target_keys_get(self):
        keys = np.zeros(max(1, self.index_sizes()[1]), np.uint32)
        self._check(self.L.mb200_target_keys_get(self.h, _p(keys)))
        return keys[:self.index_sizes()[1]]

    # ---- multi-GPU replication helpers (see bench.py: one NCCL broadcast of these buffers) ----
    def index_sizes(self):
        s = (C.c_uint64 * 4)()
        self._check(self.L.mb200_index_sizes(self.h, s))
        return [int(x) for x in s]

    def index_alloc(self, sizes):
        s = (C.c_uint64 * 4)(*sizes)
        self._check(self.L.mb200_index_alloc(self.h, s))
        self.n_targets = int(sizes[1])

    def index_buffers(self):
        ptrs = (C.c_void_p * 8)(); nb = (C.c_uint64 * 8)(); n = C.c_int(0)
        self._check(self.L.mb200_index_buffers(self.h, ptrs, nb, C.byref(n)))
        return [(int(ptrs[i] or 0), int(nb[i])) for i in range(n.value)]

    def index_commit(self):
        self._check(self.L.mb200_index_commit(self.h))

    # ---- the path's one collective, issued by the library itself (NCCL resolved at run time) ----
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self._check(self.L.mb200_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid, rank, world):
        self._check(self.L.mb200_comm_init(self.h, C.c_char_p(uid), int(rank), int(world)))

    def index_broadcast(self, root=0, comm=None):
        """-> (bytes broadcast, CUDA-event ms of the arena broadcast on this rank)"""
        nb = C.c_uint64(0); ms = C.c_double(0)
        self._check(self.L.mb200_index_broadcast(self.h, C.c_void_p(comm), int(root), C.byref(nb), C.byref(ms)))
        self.n_targets = self.index_sizes()[1]
        return int(nb.value), float(ms.value)

    def index_clone_from(self, src):
        """one process, several contexts: copy the index (and the target keys) of context `src` device to device into this one -> (bytes, wall-clock ms)"""
        nb = C.c_uint64(0); ms = C.c_double(0)
        self._check(self.L.mb200_index_clone(self.h, src.h, C.byref(nb), C.byref(ms)))
        self.n_targets = self.index_sizes()[1]
        self.target_off = getattr(src, "target_off", None)
        return int(nb.value), float(ms.value)

    # ---- host-buffer API ----
    def prefilter(self, q, qoff):
        q = np.ascontiguousarray(q, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64)
        n = len(qoff) - 1
        m = self.params.max_res_list_len
        hits = np.zeros((n, m), dtype=HIT_DTYPE); cnt = np.zeros(n, np.uint32)
        self._check(self.L.mb200_prefilter(self.h, _p(q), _p(qoff), C.c_uint32(n), _p(hits), _p(cnt)))
        return hits, cnt

    def align(self, q, qoff, hit_seq, hit_off):
        q = np.ascontiguousarray(q, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64)
        hit_seq = np.ascontiguousarray(hit_seq, np.uint32); hit_off = np.ascontiguousarray(hit_off, np.uint64)
        n = len(qoff) - 1
        alns = np.zeros(max(1, int(hit_off[-1])), dtype=ALN_DTYPE)
        if len(hit_seq) == 0:
            hit_seq = np.zeros(1, np.uint32)
        self._check(self.L.mb200_align(self.h, _p(q), _p(qoff), C.c_uint32(n), _p(hit_seq), _p(hit_off), _p(alns)))
        return alns[:int(hit_off[-1])]

    def search(self, q, qoff, copy=False):
        """-> (accepted alignments as one flat array, uint64 offsets[n+1]).  The flat array is a view of the context-owned host buffer,
        valid until the next call on this context (pass copy=True to keep it)."""
        q = np.ascontiguousarray(q, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64)
        n = len(qoff) - 1
        off = np.empty(n + 1, np.uint64); ptr = C.c_void_p()
        self._check(self.L.mb200_search(self.h, _p(q), _p(qoff), C.c_uint32(n), _p(off), C.byref(ptr)))
        return self._view(ptr, int(off[-1]), ALN_DTYPE, copy), off

    def search_resident(self, copy=False):
        """search() on the resident query batch (queries_upload / queries_from_orfs)"""
        n = self._nq
        off = np.empty(n + 1, np.uint64); ptr = C.c_void_p()
        self._check(self.L.mb200_search_resident(self.h, _p(off), C.byref(ptr)))
        return self._view(ptr, int(off[-1]), ALN_DTYPE, copy), off

    def prefilter_compact(self, q, qoff, copy=False):
        """-> (hits as one flat array, uint64 offsets[n+1]); same ownership rule as search()"""
        q = np.ascontiguousarray(q, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64)
        n = len(qoff) - 1
        off = np.empty(n + 1, np.uint64); ptr = C.c_void_p()
        self._check(self.L.mb200_prefilter_compact(self.h, _p(q), _p(qoff), C.c_uint32(n), _p(off), C.byref(ptr)))
        return self._view(ptr, int(off[-1]), HIT_DTYPE, copy), off

    @staticmethod
    def _view(ptr, count, dtype, copy):
        if count == 0:
            return np.zeros(0, dtype)
        buf = (C.c_char * (count * dtype.itemsize)).from_address(ptr.value)
        a = np.frombuffer(buf, dtype=dtype)
        return a.copy() if copy else a

    def chain(self, cands, cand_off, total_target_residues):
        cands = np.ascontiguousarray(cands, CAND_DTYPE); cand_off = np.ascontiguousarray(cand_off, np.uint64)
        nc = len(cand_off) - 1
        n = int(cand_off[-1])
        calls = np.zeros(max(1, 2 * n), dtype=CALL_DTYPE); exons = np.zeros(max(1, n), dtype=EXON_DTYPE); cnt = np.zeros(max(1, nc), np.uint32)
        self._check(self.L.mb200_chain(self.h, _p(cands), _p(cand_off), C.c_uint32(nc), C.c_uint64(total_target_residues), _p(calls), _p(cnt), _p(exons)))
        return calls, cnt[:nc], exons

    # ---- predictexons from contigs, binary records between the stages ----
    def target_keys_set(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32)
        self._check(self.L.mb200_target_keys_set(self.h, _p(keys), C.c_uint32(len(keys))))

    def predict_exons(self, nuc, off, min_len=15, max_len=32734, copy=True):
        """nuc: concatenated ASCII nucleotides, off: uint64[n_contigs+1] -> (call_off uint64[n_contigs+1], calls CALL_DTYPE, exons PEXON_DTYPE)"""
        nuc = np.ascontiguousarray(np.frombuffer(nuc, dtype=np.uint8) if isinstance(nuc, (bytes, bytearray)) else nuc, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
        nc = len(off) - 1
        call_off = np.zeros(nc + 1, np.uint64); pc = C.c_void_p(); pe = C.c_void_p(); ne = C.c_uint64(0)
        self._check(self.L.mb200_predict_exons(self.h, _p(nuc if len(nuc) else np.zeros(1, np.uint8)), _p(off), C.c_uint32(nc), int(min_len), int(max_len),
                                               _p(call_off), C.byref(pc), C.byref(pe), C.byref(ne)))
        return call_off, self._view(pc, int(call_off[-1]), CALL_DTYPE, copy), self._view(pe, int(ne.value), PEXON_DTYPE, copy)

    def text_exon_line(self, call, exon):
        buf = C.create_string_buffer(512)
        c = np.ascontiguousarray(call, CALL_DTYPE).reshape(1); e = np.ascontiguousarray(exon, PEXON_DTYPE).reshape(1)
        n = self.L.mb200_text_exon_line(_p(c), _p(e), buf, C.c_size_t(512))
        if n < 0:
            self._check(n)
        return buf.raw[:n].decode()

    def text_dp_db(self, call_off, calls, exons, contig_keys=None):
        """the dp_predictions DB of the contigs as canonical text: '>contigKey\\n' + the reference's entry, contigs without calls skipped"""
        call_off = np.ascontiguousarray(call_off, np.uint64); nc = len(call_off) - 1
        calls = np.ascontiguousarray(calls, CALL_DTYPE); exons = np.ascontiguousarray(exons, PEXON_DTYPE)
        keys = np.ascontiguousarray(contig_keys, np.uint32) if contig_keys is not None else None
        f = self.L.mb200_text_dp_db; f.restype = C.c_int64
        args = (_p(call_off), C.c_uint32(nc), _p(keys) if keys is not None else None, _p(calls if len(calls) else np.zeros(1, CALL_DTYPE)),
                _p(exons if len(exons) else np.zeros(1, PEXON_DTYPE)))
        need = f(*args, None, C.c_size_t(0))
        if need < 0:
            self._check(int(need))
        buf = C.create_string_buffer(int(need) + 1)
        f(*args, buf, C.c_size_t(int(need)))
        return buf.raw[:int(need)]

    # ---- reduceredundancy ----
    def reduce_redundancy(self, preds, pred_off, exon_keys):
        preds = np.ascontiguousarray(preds, PRED_DTYPE); pred_off = np.ascontiguousarray(pred_off, np.uint64); exon_keys = np.ascontiguousarray(exon_keys, np.uint32)
        nc = len(pred_off) - 1; n = int(pred_off[-1])
        out = np.zeros(max(1, n), PRED_OUT_DTYPE); mo = np.zeros(max(1, n), np.uint32); ro = np.zeros(max(1, n), np.uint32); cnt = np.zeros(max(1, nc), np.uint32)
        keys = exon_keys if len(exon_keys) else np.zeros(1, np.uint32)
        self._check(self.L.mb200_reduce_redundancy(self.h, _p(preds if n else np.zeros(1, PRED_DTYPE)), _p(pred_off), C.c_uint32(nc), _p(keys), C.c_uint64(len(exon_keys)),
                                                   _p(out), _p(mo), _p(ro), _p(cnt)))
        return out[:n], mo[:n], ro[:n], cnt[:nc]

    # ---- ORF extraction + translation (extractorfs + translatenucs) ----
    def extract_orfs(self, nuc, off, min_len=15, max_len=32734, start_mode=1, copy=False):
        """nuc: concatenated ASCII nucleotides (bytes / uint8 array), off: uint64[n+1] -> (records ORF_DTYPE[n_orfs], residues as a bytes-like
        uint8 array); views of context-owned buffers unless copy=True"""
        nuc = np.ascontiguousarray(np.frombuffer(nuc, dtype=np.uint8) if isinstance(nuc, (bytes, bytearray)) else nuc, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
        n = C.c_uint64(0); na = C.c_uint64(0); precs = C.c_void_p(); paa = C.c_void_p()
        self._check(self.L.mb200_extract_orfs(self.h, _p(nuc), _p(off), C.c_uint32(len(off) - 1), int(min_len), int(max_len), int(start_mode),
                                              C.byref(n), C.byref(precs), C.byref(na), C.byref(paa)))
        self._n_orfs = int(n.value)
        recs = self._view(precs, n.value, ORF_DTYPE, copy)
        aa = self._view(paa, na.value, np.dtype(np.uint8), copy)
        return recs, aa

    def queries_from_orfs(self):
        """the fragments of the last extract_orfs call become the resident query batch (no host round trip)"""
        self._check(self.L.mb200_queries_from_orfs(self.h))
        self._nq = self._n_orfs

    # ---- device-resident API (kernel-only timing) ----
    def queries_upload(self, q, qoff):
        q = np.ascontiguousarray(q, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64)
        self._nq = len(qoff) - 1
        self._check(self.L.mb200_queries_upload(self.h, _p(q), _p(qoff), C.c_uint32(self._nq)))

    def prefilter_resident(self):
        self._check(self.L.mb200_prefilter_resident(self.h))

    def align_resident(self):
        self._check(self.L.mb200_align_resident(self.h))

    def results_download(self, want_alns=False):
        n = self._nq; m = self.params.max_res_list_len
        hits = np.zeros((n, m), dtype=HIT_DTYPE); cnt = np.zeros(n, np.uint32)
        if want_alns:
            alns = np.zeros((n, m), dtype=ALN_DTYPE); acnt = np.zeros(n, np.uint32)
            self._check(self.L.mb200_results_download(self.h, _p(hits), _p(cnt), _p(alns), _p(acnt)))
            return hits, cnt, alns, acnt
        self._check(self.L.mb200_results_download(self.h, _p(hits), _p(cnt), None, None))
        return hits, cnt

    def option_set(self, name, value):
        self._check(self.L.mb200_option_set(self.h, name.encode(), C.c_int64(int(value))))

    def timer_start(self):
        """CUDA-event stopwatch on the context's stream (where every kernel of the library is launched)"""
        self._check(self.L.mb200_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double(0)
        self._check(self.L.mb200_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def query_status(self, n_queries):
        """per-query status of the last prefilter pass (0 ok, 1 too long, 2 work lists exhausted); see mb200_query_status_get"""
        st = np.zeros(n_queries, np.uint32)
        self._check(self.L.mb200_query_status_get(self.h, _p(st)))
        return st

    def profile_dump(self, path):
        self._check(self.L.mb200_profile_dump(self.h, str(path).encode()))

    def stats(self):
        s = Stats()
        self._check(self.L.mb200_stats_get(self.h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in Stats._fields_ if not k.startswith("pad")}
        d["pref_phase_cycles"] = [int(x) for x in s.pref_phase_cycles]
        return d
