"""Host-side mirror of the reference's scan interface for Python callers.

The reference exposes the scan path as `agrep [-# -i -w -x -v -n -p -I# -S# -D# -d delim -B -c] pattern file`
(agrep.c:2121-2739) and, as a library, `memagrep()/fileagrep()` (agrep.c:3282,3300).  `Pattern` takes the same
switches by name; `scan_*` return what exec() derives from the scan functions: num_of_matched and the
(lasti, print_end, j) triples handed to output() (agrep.c:3805).  All work happens in libagrepb200.so."""
import ctypes as C
from . import _lib
from ._lib import (Options, Desc, Record, Result, CorpusSpec, WANT_COUNT, WANT_RECORDS, WANT_ORDINALS, WANT_LEVELS,
                   PLAN_ALL, PLAN_ANCHORS, ENGINE_NAMES)


class AgrepError(Exception):
    pass


class Pattern:
    """agb_compile(): checksg() + preprocess() + maskgen() of the reference, plus the device plan."""

    def __init__(self, pattern, k=0, nocase=False, wordbound=False, wholeline=False, inverse=False,
                 linenum=False, ins_free=False, cost_i=0, cost_s=0, cost_d=0, bestmatch=False, delim=None):
        if isinstance(pattern, str):
            pattern = pattern.encode("latin-1")
        if isinstance(delim, str):
            delim = delim.encode("latin-1")
        self.pattern = pattern
        self.opts = Options(k=k, nocase=int(nocase), wordbound=int(wordbound), wholeline=int(wholeline),
                            inverse=int(inverse), linenum=int(linenum), ins_free=int(ins_free),
                            cost_i=cost_i, cost_s=cost_s, cost_d=cost_d, bestmatch=int(bestmatch), delim=delim)
        self._h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = _lib.lib().agb_compile(pattern, C.byref(self.opts), C.byref(self._h), err, 512)
        if rc != 0:
            raise AgrepError(err.value.decode("latin-1"))

    @property
    def desc(self):
        # a copy: the C object dies with this Pattern
        return Desc.from_buffer_copy(_lib.lib().agb_pattern_desc(self._h).contents)

    def __del__(self):
        try:
            if self._h:
                _lib.lib().agb_pattern_free(self._h)
        except Exception:
            pass

    # ---- scans -------------------------------------------------------------------------------
    def _finish(self, rc, res, recs, want):
        if rc != 0:
            raise AgrepError("agb_scan rc=%d: %s" % (rc, _lib.lib().agb_last_error().decode()))
        out = [(recs[i].begin, recs[i].end, recs[i].ordinal, recs[i].level) for i in range(res.n_records)] if recs is not None else []
        return res, out

    def scan_host(self, data, want_records=True, capacity=None, levels=False, ordinals=False):
        """data: bytes-like in host memory (the fill_buf path: H2D inside the call).
        ordinals: also fill Record.ordinal (the j that -n prints) and Result.n_closes on the device."""
        n = len(data)
        want = (WANT_RECORDS if want_records else WANT_COUNT) | (WANT_LEVELS if levels else 0) | (WANT_ORDINALS if ordinals else 0)
        cap = (capacity if capacity is not None else n // 64 + 4096) if want_records else 0
        buf = (C.c_char * n).from_buffer_copy(data) if n else None
        while True:
            recs = (Record * cap)() if cap else None
            res = Result()
            rc = _lib.lib().agb_scan_host(self._h, buf, n, want, recs, cap, C.byref(res))
            if rc != 0 or not res.truncated or capacity is not None:
                return self._finish(rc, res, recs, want)
            cap = res.n_matched          # the list did not fit (Result.truncated): once more with exactly n_matched entries

    def scan_device(self, dev_ptr, n, stream=0, d_records=0, capacity=0, levels=False, ordinals=False):
        """dev_ptr: device address of n bytes (16-byte aligned, e.g. torch tensor .data_ptr())."""
        want = (WANT_RECORDS if capacity else WANT_COUNT) | (WANT_LEVELS if levels else 0) | (WANT_ORDINALS if ordinals else 0)
        res = Result()
        rc = _lib.lib().agb_scan_device(self._h, C.c_void_p(dev_ptr), n, want, C.c_void_p(d_records), capacity,
                                        C.c_void_p(stream), C.byref(res))
        if rc != 0:
            raise AgrepError("agb_scan_device rc=%d: %s" % (rc, _lib.lib().agb_last_error().decode()))
        return res


def bestmatch_device(pattern, dev_ptr, n, stream=0, d_records=0, capacity=0, **kw):
    """The -B sweep (agrep.c:3582-3728): returns (best_k or -1, Result); with d_records/capacity (device buffer of Record)
    also the ordered list of the records at the best level."""
    if isinstance(pattern, str):
        pattern = pattern.encode("latin-1")
    d = kw.pop("delim", None)
    if isinstance(d, str):
        d = d.encode("latin-1")
    o = Options(delim=d, **{k: int(v) for k, v in kw.items()})
    res, best, err = Result(), C.c_int(-1), C.create_string_buffer(512)
    rc = _lib.lib().agb_bestmatch_device(pattern, C.byref(o), C.c_void_p(dev_ptr), n, C.c_void_p(stream),
                                        C.c_void_p(d_records), capacity, C.byref(best), C.byref(res), err, 512)
    if rc != 0:
        raise AgrepError(err.value.decode() or _lib.lib().agb_last_error().decode())
    return best.value, res


def corpus_spec(n_bytes, seed=12345, first_page=0, paragraphs=False, needle=b"", needle_every=0, needle_maxedits=0):
    if isinstance(needle, str):
        needle = needle.encode()
    return CorpusSpec(seed=seed, n_bytes=n_bytes, first_page=first_page, paragraphs=int(paragraphs),
                      needle_every=needle_every, needle=needle, needle_maxedits=needle_maxedits)


def corpus_host(n_bytes, **kw):
    """bytes of the synthetic corpus, generated by the library's host generator (same code as the device's)."""
    spec = corpus_spec(n_bytes, **kw)
    buf = C.create_string_buffer(n_bytes)
    rc = _lib.lib().agb_corpus_fill_host(C.byref(spec), buf)
    if rc != 0:
        raise AgrepError(_lib.lib().agb_last_error().decode())
    return buf.raw


def corpus_device(dev_ptr, n_bytes, stream=0, **kw):
    spec = corpus_spec(n_bytes, **kw)
    rc = _lib.lib().agb_corpus_fill_device(C.byref(spec), C.c_void_p(dev_ptr), C.c_void_p(stream))
    if rc != 0:
        raise AgrepError(_lib.lib().agb_last_error().decode())
