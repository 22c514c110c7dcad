"""srsue_b200 -- Python loader for libsrsue_gpu (the product is the C-ABI shared library built from
srsue_b200/csrc; see include/srsue_gpu/*.h).  This module only loads the library with ctypes and
offers thin helpers that pass torch CUDA tensors' device pointers through the C ABI, for the tests,
the benchmark and the smoke check.  There is no Python or CPU fallback: if the library is missing
or no GPU is usable, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRSUE_GPU_LIB", os.path.join(_HERE, "libsrsue_gpu.so"))


class Cell(C.Structure):
    # cp: 0 = normal cyclic prefix, 1 = extended (srsue_gpu_cell_t, include/srsue_gpu/srsue_gpu.h)
    _fields_ = [("nof_prb", C.c_int), ("nof_ports", C.c_int), ("cell_id", C.c_int), ("cp", C.c_int)]


class PdschCfg(C.Structure):
    _fields_ = [("sf_idx", C.c_int), ("cfi", C.c_int), ("rnti", C.c_int), ("qm", C.c_int),
                ("tbs", C.c_int), ("rv", C.c_int), ("tm", C.c_int), ("nof_prb_alloc", C.c_int),
                ("prb_mask", C.c_uint8 * 110)]


class PlanInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nfft", "nsc", "sf_len", "nof_re", "G", "C", "Kp", "Km", "Cp", "Cm", "F",
                                       "sb_cb_stride", "sb_sf_stride", "payload_stride", "max_batch")]


def host_cfo_step(cfo, nfft):
    """Phase step per sample (2^-32 turns) that removes `cfo` subcarrier spacings at nfft samples per symbol."""
    return int(lib().srsue_gpu_host_cfo_step(C.c_float(cfo), int(nfft)))


class SfDesc(C.Structure):
    """srsue_gpu_sf_desc_t (include/srsue_gpu/srsue_gpu.h)"""
    _fields_ = [("cell", Cell), ("cfg", PdschCfg), ("iq", C.c_void_p), ("payload", C.c_void_p),
                ("softbuffer_id", C.c_int64), ("new_data", C.c_int32), ("crc_ok", C.c_int32), ("n_iter", C.c_int32),
                ("meas", C.c_float * 5), ("cfo", C.c_float)]


class SyncResult(C.Structure):
    """srsue_gpu_sync_result_t"""
    _fields_ = [("peak_pos", C.c_int32), ("n_id_2", C.c_int32), ("n_id_1", C.c_int32), ("sf5", C.c_int32), ("valid", C.c_int32),
                ("peak", C.c_float), ("mean_power", C.c_float), ("cfo", C.c_float), ("sss_corr", C.c_float), ("cp", C.c_int32)]


class GpuError(RuntimeError):
    pass


_lib = None


def build(verbose=False):
    """Compile libsrsue_gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    import subprocess
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")], stdout=out)
    return LIB_PATH


def lib():
    """Load libsrsue_gpu.so (raises if it has not been built -- there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GpuError("libsrsue_gpu.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.srsue_gpu_last_error.restype = C.c_char_p
        L.srsue_gpu_host_alloc.restype = C.c_void_p
        L.srsue_gpu_host_alloc.argtypes = [C.c_uint64]
        L.srsue_gpu_host_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise GpuError("%s failed (%d): %s" % (what, rc, lib().srsue_gpu_last_error().decode()))


def _ptr(t):
    """device (or host) pointer of a torch tensor / numpy array / None as c_void_p"""
    if t is None:
        return C.c_void_p(0)
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Context:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(lib().srsue_gpu_ctx_create(C.byref(self.h), device), "srsue_gpu_ctx_create")
        self.device = device

    def close(self):
        if self.h:
            lib().srsue_gpu_ctx_destroy(self.h)
            self.h = C.c_void_p()

    # ---- turbo ----
    @staticmethod
    def tdec_geometry(K):
        W, P, n = C.c_int(), C.c_int(), C.c_int()
        _check(lib().srsue_gpu_tdec_geometry(K, C.byref(W), C.byref(P), C.byref(n)), "srsue_gpu_tdec_geometry")
        return W.value, P.value, n.value

    def tdec_import(self, d_triples, n_cb, K, d_tcb):
        _check(lib().srsue_gpu_tdec_import(self.h, _ptr(d_triples), n_cb, K, _ptr(d_tcb), _stream()), "tdec_import")

    def tdec_export(self, d_tcb, n_cb, K, d_triples):
        _check(lib().srsue_gpu_tdec_export(self.h, _ptr(d_tcb), n_cb, K, _ptr(d_triples), _stream()), "tdec_export")

    def tdec_decode(self, d_tcb, n_cb, K, max_iter, crc_type, d_bits, d_status):
        _check(lib().srsue_gpu_tdec_decode(self.h, _ptr(d_tcb), n_cb, K, max_iter, crc_type, _ptr(d_bits), _ptr(d_status),
                                           _stream()), "tdec_decode")

    def tdec_run_all(self, d_triples, n_cb, K, max_iter, crc_type, d_bits, d_status):
        _check(lib().srsue_gpu_tdec_run_all(self.h, _ptr(d_triples), n_cb, K, max_iter, crc_type, _ptr(d_bits),
                                            _ptr(d_status), _stream()), "tdec_run_all")

    def tdec_run_all_host(self, h_triples, n_cb, K, max_iter, crc_type, h_bits, h_status):
        _check(lib().srsue_gpu_tdec_run_all_host(self.h, _ptr(h_triples), n_cb, K, max_iter, crc_type, _ptr(h_bits),
                                                 _ptr(h_status)), "tdec_run_all_host")

    def cell_search(self, d_iq, n_bufs, n_samples, stride, d_result, force_n_id_2=-1, first_pos=0, nfft=128, cp_mode=0):
        """cp_mode 0: SSS behind a normal cyclic prefix, 1: extended, 2: both, the better one reported in result.cp"""
        lib().srsue_gpu_cell_search_cp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int,
                                                   C.c_void_p, C.c_void_p]
        _check(lib().srsue_gpu_cell_search_cp(self.h, _ptr(d_iq), n_bufs, n_samples, stride, nfft, force_n_id_2, first_pos, cp_mode,
                                              _ptr(d_result), _stream()),
               "cell_search")

    def tdec_last_launch(self):
        g, b, s, n = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib().srsue_gpu_tdec_last_launch(self.h, C.byref(g), C.byref(b), C.byref(s), C.byref(n))
        return dict(grid=g.value, block=b.value, smem=s.value, cb_per_cta=n.value)

    def last_launch_count(self):
        return lib().srsue_gpu_last_launch_count(self.h)


class PdschPlan:
    def __init__(self, ctx, cell, cfg, max_batch):
        self.ctx = ctx
        self.h = C.c_void_p()
        _check(lib().srsue_gpu_pdsch_plan_create(ctx.h, C.byref(cell), C.byref(cfg), max_batch, C.byref(self.h)),
               "srsue_gpu_pdsch_plan_create")
        self.info = PlanInfo()
        lib().srsue_gpu_pdsch_plan_info(self.h, C.byref(self.info))

    def close(self):
        if self.h:
            lib().srsue_gpu_pdsch_plan_destroy(self.h)
            self.h = C.c_void_p()

    def ofdm_rx(self, n_sf, d_iq, d_sf, d_cfo_steps=None, cfo_step=0):
        """OFDM demodulation; with d_cfo_steps (int32 per subframe) or cfo_step (all subframes) the carrier-offset
        rotation of SPEC.md 14 is applied as the samples are loaded (steps from host_cfo_step)."""
        if d_cfo_steps is None and cfo_step == 0:
            _check(lib().srsue_gpu_ofdm_rx(self.h, n_sf, _ptr(d_iq), _ptr(d_sf), _stream()), "ofdm_rx")
        else:
            _check(lib().srsue_gpu_ofdm_rx_cfo(self.h, n_sf, _ptr(d_iq), _ptr(d_sf), _ptr(d_cfo_steps), C.c_int32(cfo_step),
                                               _stream()), "ofdm_rx_cfo")

    def ofdm_rx_sc16(self, n_sf, d_iq16, scale, d_sf, d_cfo_steps=None, cfo_step=0):
        """OFDM demodulation of int16 {re, im} samples; sample = float(v) * scale, then the optional carrier-offset rotation."""
        _check(lib().srsue_gpu_ofdm_rx_sc16_cfo(self.h, n_sf, _ptr(d_iq16), C.c_float(scale), _ptr(d_sf), _ptr(d_cfo_steps),
                                                C.c_int32(cfo_step), _stream()), "ofdm_rx_sc16")

    def set_cfo(self, d_cfo_steps=None, cfo_step=0):
        """carrier-offset correction for decode_batch / decode_batch_host of this plan; (None, 0) switches it off"""
        _check(lib().srsue_gpu_pdsch_plan_set_cfo(self.h, _ptr(d_cfo_steps), C.c_int32(cfo_step)), "set_cfo")

    def set_min_iter(self, min_iter):
        """a passing code-block CRC ends the block only from this turbo iteration on (1 = default, max_iter = fixed count)"""
        _check(lib().srsue_gpu_pdsch_plan_set_min_iter(self.h, int(min_iter)), "set_min_iter")

    def set_iq_format(self, sc16, scale=1.0 / 32768.0):
        """decode_batch / decode_batch_host take int16 {re, im} samples (sc16=True) or complex64 (False, the default)."""
        _check(lib().srsue_gpu_pdsch_plan_set_iq_format(self.h, 1 if sc16 else 0, C.c_float(scale)), "set_iq_format")

    def chest(self, n_sf, d_sf, d_ce, d_meas):
        _check(lib().srsue_gpu_chest(self.h, n_sf, _ptr(d_sf), _ptr(d_ce), _ptr(d_meas), _stream()), "chest")

    def pcfich_decode(self, n_sf, d_sf, d_ce, d_meas, noise_est, noise_mode, d_cfi, d_corr=None):
        _check(lib().srsue_gpu_pcfich_decode(self.h, n_sf, _ptr(d_sf), _ptr(d_ce), _ptr(d_meas), C.c_float(noise_est), noise_mode,
                                             _ptr(d_cfi), _ptr(d_corr), _stream()), "pcfich_decode")

    def pbch_decode(self, n_sf, d_sf, d_ce, d_meas, noise_est, noise_mode, d_result, d_mib):
        _check(lib().srsue_gpu_pbch_decode(self.h, n_sf, _ptr(d_sf), _ptr(d_ce), _ptr(d_meas), C.c_float(noise_est), noise_mode,
                                           _ptr(d_result), _ptr(d_mib), _stream()), "pbch_decode")

    def phich_decode(self, n_sf, d_sf, d_ce, d_meas, noise_est, noise_mode, n_group, n_seq, d_ack, d_metric=None, ng_x6=6):
        _check(lib().srsue_gpu_phich_decode(self.h, n_sf, _ptr(d_sf), _ptr(d_ce), _ptr(d_meas), C.c_float(noise_est), noise_mode,
                                            ng_x6, n_group, n_seq, _ptr(d_ack), _ptr(d_metric), _stream()), "phich_decode")

    def pdcch_info(self, ng_x6=6):
        a, b = C.c_int(), C.c_int()
        _check(lib().srsue_gpu_pdcch_info(self.h, ng_x6, C.byref(a), C.byref(b)), "pdcch_info")
        return a.value, b.value

    def pdcch_extract_llr(self, n_sf, d_sf, d_ce, d_meas, noise_est, noise_mode, d_llr, ng_x6=6):
        _check(lib().srsue_gpu_pdcch_extract_llr(self.h, n_sf, _ptr(d_sf), _ptr(d_ce), _ptr(d_meas), C.c_float(noise_est), noise_mode,
                                                 ng_x6, _ptr(d_llr), _stream()), "pdcch_extract_llr")

    def pdcch_find_dci(self, n_sf, d_llr, rnti, nof_bits, d_found, d_bits, d_rem=None, common=0, ng_x6=6, first_bit=-1):
        n = lib().srsue_gpu_pdcch_find_dci(self.h, n_sf, _ptr(d_llr), ng_x6, rnti, common, nof_bits, first_bit, _ptr(d_found), _ptr(d_bits),
                                           _ptr(d_rem), _stream())
        if n < 0:
            _check(n, "pdcch_find_dci")
        return n

    def pdsch_llr(self, n_sf, d_sf, d_ce, d_meas, noise_est, noise_mode, accumulate, d_softbuf, d_dbg_d=None, d_dbg_e=None):
        _check(lib().srsue_gpu_pdsch_llr(self.h, n_sf, _ptr(d_sf), _ptr(d_ce), _ptr(d_meas), C.c_float(noise_est), noise_mode,
                                         accumulate, _ptr(d_softbuf), _ptr(d_dbg_d), _ptr(d_dbg_e), _stream()), "pdsch_llr")

    def chest_pilots(self, n_sf, d_sf, d_pilots, d_meas):
        _check(lib().srsue_gpu_chest_pilots(self.h, n_sf, _ptr(d_sf), _ptr(d_pilots), _ptr(d_meas), _stream()), "chest_pilots")

    def pdsch_llr_fused(self, n_sf, d_sf, d_pilots, d_meas, noise_est, noise_mode, accumulate, d_softbuf):
        _check(lib().srsue_gpu_pdsch_llr_fused(self.h, n_sf, _ptr(d_sf), _ptr(d_pilots), _ptr(d_meas), C.c_float(noise_est),
                                               noise_mode, accumulate, _ptr(d_softbuf), _stream()), "pdsch_llr_fused")

    def pdsch_turbo(self, n_sf, d_softbuf, max_iter, d_payload, d_tb_status, d_cb_status=None):
        _check(lib().srsue_gpu_pdsch_turbo(self.h, n_sf, _ptr(d_softbuf), max_iter, _ptr(d_payload), _ptr(d_tb_status),
                                           _ptr(d_cb_status), _stream()), "pdsch_turbo")

    def decode_batch(self, n_sf, d_iq, noise_est, noise_mode, max_iter, d_payload, d_tb_status, d_softbuf=None,
                     accumulate=0, d_meas=None):
        _check(lib().srsue_gpu_pdsch_decode_batch(self.h, n_sf, _ptr(d_iq), C.c_float(noise_est), noise_mode, max_iter,
                                                  accumulate, _ptr(d_softbuf), _ptr(d_payload), _ptr(d_tb_status),
                                                  _ptr(d_meas), _stream()), "pdsch_decode_batch")

    def decode_batch_host(self, n_sf, h_iq, noise_est, noise_mode, max_iter, h_payload, h_tb_status, h_meas=None):
        _check(lib().srsue_gpu_pdsch_decode_batch_host(self.h, n_sf, _ptr(h_iq), C.c_float(noise_est), noise_mode, max_iter,
                                                       _ptr(h_payload), _ptr(h_tb_status), _ptr(h_meas)),
               "pdsch_decode_batch_host")


class UlschCfg(C.Structure):
    """srsue_gpu_ulsch_cfg_t"""
    _fields_ = [(n, C.c_int) for n in ("tbs", "qm", "nof_prb", "n_symb", "rv", "rnti", "sf_idx", "cell_id")]


class UlschPlan:
    """Uplink shared-channel encoder (srsue_gpu_ulsch_*): transport blocks -> interleaved, scrambled PUSCH bits."""

    def __init__(self, ctx, tbs, qm, nof_prb, rv=0, rnti=0x1234, sf_idx=0, cell_id=1, n_symb=12, max_batch=64):
        self.ctx = ctx
        self.cfg = UlschCfg(tbs, qm, nof_prb, n_symb, rv, rnti, sf_idx, cell_id)
        self.h = C.c_void_p()
        _check(lib().srsue_gpu_ulsch_plan_create(ctx.h, C.byref(self.cfg), max_batch, C.byref(self.h)), "srsue_gpu_ulsch_plan_create")
        G, Cb, Kp, Km = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().srsue_gpu_ulsch_plan_info(self.h, C.byref(G), C.byref(Cb), C.byref(Kp), C.byref(Km)), "srsue_gpu_ulsch_plan_info")
        self.G, self.C, self.Kp, self.Km = G.value, Cb.value, Kp.value, Km.value

    def close(self):
        if self.h:
            lib().srsue_gpu_ulsch_plan_destroy.argtypes = [C.c_void_p]
            lib().srsue_gpu_ulsch_plan_destroy(self.h)
            self.h = C.c_void_p()

    def encode(self, n_tb, d_payload, d_bits):
        _check(lib().srsue_gpu_ulsch_encode(self.h, n_tb, _ptr(d_payload), _ptr(d_bits), _stream()), "srsue_gpu_ulsch_encode")

    def encode_host(self, n_tb, h_payload, h_bits):
        _check(lib().srsue_gpu_ulsch_encode_host(self.h, n_tb, _ptr(h_payload), _ptr(h_bits)), "srsue_gpu_ulsch_encode_host")


class Batch:
    """srsue_gpu_batch_*: heterogeneous subframe streams.  Keeps the numpy buffers alive until wait()."""

    def __init__(self, ctx, max_subframes, noise_est=0.01, noise_mode=0, max_iter=4, devices=None):
        """ctx: a Context (one device), or None with devices=[...]: srsue_gpu_batch_create_multi -- one handle over several
        GPUs, the library owns a context, a single-device batch and a host thread per device"""
        self.ctx = ctx
        self.h = C.c_void_p()
        lib().srsue_gpu_batch_softbuffer_release.argtypes = [C.c_void_p, C.c_int64]
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            _check(lib().srsue_gpu_batch_create_multi(arr, len(devices), max_subframes, C.c_float(noise_est), noise_mode, max_iter,
                                                      C.byref(self.h)), "srsue_gpu_batch_create_multi")
        else:
            _check(lib().srsue_gpu_batch_create(ctx.h, max_subframes, C.c_float(noise_est), noise_mode, max_iter, C.byref(self.h)),
                   "srsue_gpu_batch_create")
        self._keep = None

    def device_shares(self):
        """multi-GPU handles: [(subframes, estimated work)] per device of the last submission"""
        n = C.c_int()
        sf = (C.c_int * 64)()
        w = (C.c_double * 64)()
        lib().srsue_gpu_batch_device_shares(self.h, C.byref(n), sf, w, 64)
        return [(sf[i], w[i]) for i in range(n.value)]

    @staticmethod
    def prepare(items):
        """items: list of dicts {cell, cfg, iq (complex64 array; int16 pairs after set_iq_format(True)), softbuffer_id (default -1), new_data (default 1),
        payload (optional uint8 array), cfo (subcarrier spacings, default 0)} -> (descriptor array, payload arrays, items), reusable across submissions"""
        import numpy as np
        n = len(items)
        descs = (SfDesc * n)()
        payloads = []
        for d, it in zip(descs, items):
            d.cell, d.cfg = it["cell"], it["cfg"]
            pl = it.get("payload")
            if pl is None:
                pl = np.zeros((it["cfg"].tbs + 7) // 8, np.uint8)
            payloads.append(pl)
            d.iq, d.payload = it["iq"].ctypes.data, pl.ctypes.data
            d.softbuffer_id = it.get("softbuffer_id", -1)
            d.new_data = it.get("new_data", 1)
            d.cfo = it.get("cfo", 0.0)
        return descs, payloads, items

    def set_iq_format(self, sc16, scale=1.0 / 32768.0):
        """every item's iq is an int16 {re, im} array (sc16=True; sample = float(v) * scale) or complex64 (False, default)"""
        _check(lib().srsue_gpu_batch_set_iq_format(self.h, 1 if sc16 else 0, C.c_float(scale)), "srsue_gpu_batch_set_iq_format")

    def submit_prepared(self, prepared):
        self._keep = None
        _check(lib().srsue_gpu_batch_submit(self.h, prepared[0], len(prepared[0])), "srsue_gpu_batch_submit")
        self._keep = prepared                 # a refused submission leaves nothing to wait for

    def submit(self, items):
        self.submit_prepared(self.prepare(items))

    def submit_blind(self, items, ng_x6=6, payload_cap=None):
        """srsue_gpu_batch_submit_blind: items carry cell, cfg (only sf_idx and rnti are read), iq and a payload buffer of
        payload_cap bytes; CFI, grant and transport block come back (wait() results gain 'cfg')"""
        import numpy as np
        if payload_cap is None:
            payload_cap = 75376 // 8
        for it in items:
            it.setdefault("payload", np.zeros(payload_cap, np.uint8))
        prepared = self.prepare(items)
        self._keep = None
        _check(lib().srsue_gpu_batch_submit_blind(self.h, prepared[0], len(prepared[0]), ng_x6, payload_cap), "srsue_gpu_batch_submit_blind")
        self._keep = prepared
        self._blind = True

    def wait(self):
        _check(lib().srsue_gpu_batch_wait(self.h), "srsue_gpu_batch_wait")
        if self._keep is None:
            return []
        descs, payloads, _ = self._keep
        self._keep = None
        if getattr(self, "_blind", False):
            self._blind = False
            return [dict(payload=pl[:(d.cfg.tbs + 7) // 8], crc_ok=d.crc_ok, n_iter=d.n_iter, meas=list(d.meas), cfg=d.cfg,
                         tbs=d.cfg.tbs, cfi=d.cfg.cfi) for d, pl in zip(descs, payloads)]
        return [dict(payload=pl, crc_ok=d.crc_ok, n_iter=d.n_iter, meas=list(d.meas)) for d, pl in zip(descs, payloads)]

    def release_softbuffer(self, sid):
        _check(lib().srsue_gpu_batch_softbuffer_release(self.h, sid), "srsue_gpu_batch_softbuffer_release")

    def stats(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().srsue_gpu_batch_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(plans=a.value, softbuffers=b.value, launches=c.value)

    def close(self):
        if self.h:
            lib().srsue_gpu_batch_destroy(self.h)
            self.h = C.c_void_p()


def make_cell(nof_prb, nof_ports=1, cell_id=1, cp=0):
    return Cell(nof_prb, nof_ports, cell_id, cp)


def make_cfg(cell, sf_idx=1, cfi=1, rnti=0x1234, qm=2, tbs=152, rv=0, tm=1, prbs=None, prbs_slot1=None):
    """prbs: allocated PRBs (all when None); prbs_slot1: the PRBs of the second slot when they differ from the first
    (distributed virtual resource blocks) -- prb_mask then holds 1 (both slots), 2 (slot 0 only) or 4 (slot 1 only)"""
    cfg = PdschCfg()
    cfg.sf_idx, cfg.cfi, cfg.rnti, cfg.qm, cfg.tbs, cfg.rv, cfg.tm = sf_idx, cfi, rnti, qm, tbs, rv, tm
    s0 = set(range(cell.nof_prb) if prbs is None else prbs)
    s1 = s0 if prbs_slot1 is None else set(prbs_slot1)
    for p in s0 | s1:
        cfg.prb_mask[p] = 1 if (p in s0 and p in s1) else 2 if p in s0 else 4
    cfg.nof_prb_alloc = len(s0)
    return cfg
