"""CPU: include/b200fe.h, the ctypes binding (lighting-asr_b200/_lib.py) and the stubs printed in INTEGRATION.md declare the
SAME structs (field names, order, C types) and the library exports every function the header declares."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "b200fe.h")).read()

_CT = {"int": C.c_int, "unsigned int": C.c_uint, "float": C.c_float, "long long": C.c_longlong, "unsigned long long": C.c_ulonglong,
       "double": C.c_double}


def _strip_comments(txt):
    return re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)


def parse_struct(name):
    """[(field, ctypes type)] of `typedef struct name { ... } name;` in the header."""
    m = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), _strip_comments(HEADER), flags=re.S)
    assert m, name
    fields = []
    for decl in m.group(1).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        first, *more = [d.strip() for d in decl.split(",")]
        mm = re.match(r"(.*?)(\**)\s*(\w+)$", first)
        base, stars, fname = mm.group(1).strip(), mm.group(2), mm.group(3)
        base = base.replace("const ", "").strip()
        names = [(fname, bool(stars))] + [(x.lstrip("*").strip(), x.startswith("*")) for x in more]
        for fn, ptr in names:
            fields.append((fn, C.c_void_p if ptr else _CT[base]))
    return fields


def test_structs_match_the_ctypes_binding(lasr_b200):
    L = lasr_b200._lib
    for cname, cls in (("b200fe_opts", L.Opts), ("b200fe_fbank_args", L.FbankArgs), ("b200fe_post_args", L.PostArgs), ("b200fe_warp_args", L.WarpArgs)):
        want = parse_struct(cname)
        got = [(n, t) for n, t in cls._fields_]
        assert [n for n, _ in got] == [n for n, _ in want], cname
        for (n, tg), (_, tw) in zip(got, want):
            assert C.sizeof(tg) == C.sizeof(tw) and (tg is C.c_void_p) == (tw is C.c_void_p), (cname, n)
        ref = type("Ref_" + cname, (C.Structure,), {"_fields_": want})
        assert C.sizeof(ref) == C.sizeof(cls), cname
    a = L.FbankArgs()
    assert a.struct_size == C.sizeof(L.FbankArgs) and L.PostArgs().struct_size == C.sizeof(L.PostArgs) and L.WarpArgs().struct_size == C.sizeof(L.WarpArgs)


def test_library_exports_every_declared_function(lasr_b200):
    lib = lasr_b200._lib.load()
    declared = set(re.findall(r"\b(b200fe_\w+)\s*\(", _strip_comments(HEADER)))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), "declared in include/b200fe.h but not exported: " + name
    assert set(lasr_b200._lib.EXPORTS) <= declared
    missing = declared - set(lasr_b200._lib.EXPORTS)
    assert not missing, "declared but not bound in _lib.EXPORTS: %s" % sorted(missing)


def test_stale_struct_is_rejected_without_a_gpu(lasr_b200):
    """struct_size is checked before anything touches the device: a binding built against an older header fails loudly."""
    L = lasr_b200._lib
    lib = L.load()
    a = L.FbankArgs()
    a.struct_size -= 8
    assert lib.b200fe_fbank_fused(C.c_void_p(1), C.byref(a), None) == -1
    assert b"struct_size" in lib.b200fe_last_error()
    q = L.PostArgs()
    q.struct_size = 0
    assert lib.b200fe_postpass(C.c_void_p(1), C.byref(q), None) == -1
    w = L.WarpArgs()
    w.struct_size += 4
    assert lib.b200fe_time_warp(C.c_void_p(1), C.byref(w), None) == -1


def test_integration_md_prints_the_current_structs(lasr_b200):
    """INTEGRATION.md shows the reference-side ctypes stubs: every field of every struct appears there, in order."""
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for cname in ("b200fe_opts", "b200fe_fbank_args", "b200fe_post_args", "b200fe_warp_args"):
        pos = txt.find("# " + cname)
        assert pos >= 0, "INTEGRATION.md lacks the stub of " + cname
        for fname, _ in parse_struct(cname):
            nxt = txt.find('"%s"' % fname, pos)
            assert nxt >= 0, "INTEGRATION.md: field %s of %s missing or out of order (run tools/gen_integration_structs.py)" % (fname, cname)
            pos = nxt
