"""The Rust binding `rust/rt-sys/src/lib.rs` cannot be compiled in this image (no cargo/rustc).  This test keeps it from drifting:
every `#[repr(C)]` struct, every `extern "C"` prototype and every constant is parsed out of BOTH `include/rt_api.h` and `lib.rs`
and compared -- field order, names, types; function names, argument order and types, return types; enum values.  The struct
sizes asserted in `lib.rs` (and by `static_assert` in csrc/rt_api.cu) are re-derived from the header with ctypes.

The seam being bound: `render_scene` (`/root/reference/src/rendering.rs:21`, call site `src/main.rs:55`)."""
import ctypes as C
import os
import re

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "rt_api.h")
LIB_RS = os.path.join(ROOT, "rust", "rt-sys", "src", "lib.rs")

C_TO_RUST = {"int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "uint32_t": "u32", "uint8_t": "u8", "double": "f64", "float": "f32", "int": "c_int", "char": "c_char",
             "void": "c_void"}
CTYPES = {"i32": C.c_int32, "i64": C.c_int64, "u64": C.c_uint64, "f64": C.c_double, "f32": C.c_float}


def _strip_c(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _c_type(decl):
    """'const double*' -> '*const f64', 'RtScene**' -> '*mut *mut RtScene', 'int32_t' -> 'i32'."""
    decl = decl.strip()
    stars = decl.count("*")
    base = decl.replace("*", " ").split()
    const = "const" in base
    base = [b for b in base if b != "const"]
    name = C_TO_RUST.get(base[0], base[0])
    if stars == 0:
        return name
    if decl.replace(" ", "").endswith("*const*"):              # `RtScene* const*`: const pointer array of mutable scenes
        return "*const *mut " + name
    out = name
    for k in range(stars):
        out = ("*const " if (const and k == 0) else "*mut ") + out
    return out


def parse_header():
    text = _strip_c(open(HEADER).read())
    structs = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", text, flags=re.S):
        fields = []
        for stmt in m.group(2).split(";"):
            stmt = " ".join(stmt.split())
            if not stmt:
                continue
            mm = re.match(r"(.*?)([\w\[\], ]+)$", stmt)
            # split "type a, b[3], c" into type + declarators
            tm = re.match(r"((?:const )?\w+\s*\**)\s*(.*)", stmt)
            ctype, decls = tm.group(1).strip(), tm.group(2)
            for d in decls.split(","):
                d = d.strip()
                arr = re.match(r"(\w+)\[(\d+)\]", d)
                if arr:
                    fields.append((arr.group(1), f"[{_c_type(ctype)}; {arr.group(2)}]"))
                else:
                    stars = d.count("*")
                    fields.append((d.replace("*", "").strip(), _c_type(ctype + "*" * stars)))
        structs[m.group(1)] = fields
    funcs = {}
    for m in re.finditer(r"\n\s*((?:const )?\w+\s*\**)\s*(rt_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        alist = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                am = re.match(r"(.*?)(\w+)$", a)
                alist.append((am.group(2), _c_type(am.group(1))))
        funcs[name] = (None if ret == "void" else _c_type(ret), alist)
    consts = {}
    for m in re.finditer(r"enum \{(.*?)\};", text, flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = item.split("=")
                nxt = int(v.strip())
                consts[k.strip()] = nxt
            else:
                consts[item] = nxt
            nxt += 1
    mv = re.search(r"#define RT_API_VERSION (\d+)", text)
    consts["RT_API_VERSION"] = int(mv.group(1))
    return structs, funcs, consts


def parse_rust():
    text = re.sub(r"//[^\n]*", " ", open(LIB_RS).read())
    structs = {}
    for m in re.finditer(r"#\[repr\(C\)\](?:\s*#\[derive\([^)]*\)\])?\s*pub struct (\w+) \{(.*?)\n\}", text, flags=re.S):
        fields = []
        for f in m.group(2).split(",\n"):
            f = " ".join(f.split()).rstrip(",")
            if not f:
                continue
            fm = re.match(r"(?:pub )?(\w+): (.*)", f)
            fields.append((fm.group(1), fm.group(2).strip()))
        structs[m.group(1)] = fields
    funcs = {}
    ext = re.search(r'extern "C" \{(.*?)\n\}', text, flags=re.S).group(1)
    for m in re.finditer(r"pub fn (\w+)\((.*?)\)(?: -> ([^;]+))?;", ext, flags=re.S):
        alist = []
        for a in m.group(2).split(","):
            a = " ".join(a.split())
            if not a:
                continue
            k, t = a.split(": ", 1)
            alist.append((k.replace("r#", ""), t.strip()))
        funcs[m.group(1)] = (m.group(3).strip() if m.group(3) else None, alist)
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (\w+): c_int = (\d+);", text)}
    sizes = {m.group(1): int(m.group(2)) for m in re.finditer(r"size_of::<(\w+)>\(\) == (\d+)", text)}
    offsets = {(m.group(1), m.group(2)): int(m.group(3)) for m in re.finditer(r"offset_of!\((\w+), (\w+)\) == (\d+)", text)}
    return structs, funcs, consts, sizes, offsets


def test_repr_c_structs_match_the_header():
    hs, _, _ = parse_header()
    rs, _, _, _, _ = parse_rust()
    assert set(hs) == {"RtSceneDesc", "RtSceneDesc2", "RtSceneInfo", "RtRenderParams", "RtStats"}
    for name, fields in hs.items():
        assert name in rs, f"{name} missing from lib.rs"
        assert rs[name] == fields, (name, [x for x in zip(fields, rs[name]) if x[0] != x[1]][:3], len(fields), len(rs[name]))
    assert rs["RtScene"] == [("_private", "[u8; 0]")]                       # opaque handle


def test_extern_c_prototypes_match_the_header():
    _, hf, _ = parse_header()
    _, rf, _, _, _ = parse_rust()
    assert len(hf) >= 27
    assert set(hf) == set(rf), (sorted(set(hf) - set(rf)), sorted(set(rf) - set(hf)))
    for name, (ret, args) in hf.items():
        rret, rargs = rf[name]
        assert rret == ret, (name, ret, rret)
        assert [t for _, t in rargs] == [t for _, t in args], (name, args, rargs)
        assert [k for k, _ in rargs] == [k for k, _ in args], (name, args, rargs)


def test_constants_match_the_header():
    _, _, hc = parse_header()
    _, _, rc, _, _ = parse_rust()
    assert hc == rc, {k: (hc.get(k), rc.get(k)) for k in set(hc) | set(rc) if hc.get(k) != rc.get(k)}


def _ctypes_struct(name, structs, cache):
    if name in cache:
        return cache[name]
    fields = []
    for k, t in structs[name]:
        arr = re.match(r"\[(\w+); (\d+)\]", t)
        if arr:
            ct = CTYPES[arr.group(1)] * int(arr.group(2))
        elif t.startswith("*"):
            ct = C.c_void_p
        elif t in CTYPES:
            ct = CTYPES[t]
        else:
            ct = _ctypes_struct(t, structs, cache)
        fields.append((k, ct))
    cache[name] = type(name, (C.Structure,), {"_fields_": fields})
    return cache[name]


def test_sizes_and_offsets_asserted_in_rust_follow_from_the_header(rt):
    hs, _, _ = parse_header()
    _, _, _, sizes, offsets = parse_rust()
    cache = {}
    assert set(sizes) == set(hs)
    for name, size in sizes.items():
        assert C.sizeof(_ctypes_struct(name, hs, cache)) == size, name
    assert len(offsets) >= 10
    for (name, field), off in offsets.items():
        assert getattr(_ctypes_struct(name, hs, cache), field).offset == off, (name, field)
    # the ctypes mirrors the tests and bench.py use have the same sizes
    for name in ("RtSceneDesc", "RtSceneDesc2", "RtSceneInfo", "RtRenderParams", "RtStats"):
        assert C.sizeof(getattr(rt, name)) == sizes[name], name
    # and the C side asserts the same numbers at compile time
    src = open(os.path.join(ROOT, "raytracing-course-2024_b200", "csrc", "rt_api.cu")).read()
    for name, size in sizes.items():
        assert f"sizeof({name}) == {size}" in src, name
