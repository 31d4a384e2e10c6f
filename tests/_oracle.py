"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE (the checker, never the product)."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


class Opts(C.Structure):
    _fields_ = [("k", C.c_int), ("nocase", C.c_int), ("wordbound", C.c_int), ("wholeline", C.c_int),
                ("inverse", C.c_int), ("linenum", C.c_int), ("ins_free", C.c_int),
                ("cost_i", C.c_int), ("cost_s", C.c_int), ("cost_d", C.c_int),
                ("bestmatch", C.c_int), ("width", C.c_int), ("delim", C.c_char_p)]


class Automaton(C.Structure):
    _fields_ = [("mask", C.c_uint64 * 256), ("init0", C.c_uint64), ("init1", C.c_uint64),
                ("noerr", C.c_uint64), ("endpos", C.c_uint64), ("dendpos", C.c_uint64),
                ("dmask", C.c_uint64), ("wildmask", C.c_uint64), ("M", C.c_int), ("L", C.c_int),
                ("dpat", C.c_ubyte * 18), ("and_mode", C.c_int), ("user_delim", C.c_int),
                ("outtail", C.c_int), ("sgrep", C.c_int), ("engine", C.c_int),
                ("k", C.c_int), ("inverse", C.c_int), ("jump", C.c_int),
                ("ci", C.c_int), ("cs", C.c_int), ("cd", C.c_int), ("lut_fold", C.c_int),
                ("lit", C.c_ubyte * 256), ("litlen", C.c_int), ("lit_word", C.c_int)]


class Record(C.Structure):
    _fields_ = [("begin", C.c_int64), ("end", C.c_int64), ("ordinal", C.c_int64), ("level", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        _lib.orc_compile.argtypes = [C.c_char_p, C.POINTER(Opts), C.POINTER(Automaton), C.c_char_p, C.c_size_t]
        _lib.orc_scan.restype = C.c_int64
        _lib.orc_scan.argtypes = [C.POINTER(Automaton), C.c_char_p, C.c_uint64, C.POINTER(Record), C.c_uint64]
        _lib.orc_scan_levels.restype = C.c_int64
        _lib.orc_scan_levels.argtypes = [C.POINTER(Automaton), C.c_int, C.c_char_p, C.c_uint64,
                                         C.POINTER(C.c_uint64), C.POINTER(Record), C.c_uint64, C.c_int]
    return _lib


class OracleError(Exception):
    pass


def compile(pattern, **kw):
    if isinstance(pattern, str):
        pattern = pattern.encode("latin-1")
    o = Opts()
    for k, v in kw.items():
        if k == "delim" and isinstance(v, str):
            v = v.encode("latin-1")
        setattr(o, k, v)
    a = Automaton()
    err = C.create_string_buffer(256)
    if lib().orc_compile(pattern, C.byref(o), C.byref(a), err, 256) != 0:
        raise OracleError(err.value.decode())
    return a


def scan(a, text, want_records=True, cap=None):
    """returns (count, [(begin, end, ordinal), ...])"""
    n = len(text)
    if not want_records:
        return lib().orc_scan(C.byref(a), text, n, None, 0), []
    cap = cap or (n + 2)          # an empty record is a record: up to one per text byte
    recs = (Record * cap)()
    cnt = lib().orc_scan(C.byref(a), text, n, recs, cap)
    return cnt, [(recs[i].begin, recs[i].end, recs[i].ordinal) for i in range(min(cnt, cap))]


def scan_levels(a, kmax, text, want_level=-1, cap=None):
    n = len(text)
    cap = cap or (n + 2)          # an empty record is a record: up to one per text byte
    recs = (Record * cap)()
    hist = (C.c_uint64 * 9)()
    cnt = lib().orc_scan_levels(C.byref(a), kmax, text, n, hist, recs, cap, want_level)
    return cnt, list(hist), [(recs[i].begin, recs[i].end, recs[i].ordinal, recs[i].level) for i in range(min(cnt, cap))]
