"""The Rust shim (gsv-cuda/src/ffi.rs) against the C header (include/gsv_cuda.h): every exported function with the
same argument list, every struct with the same fields in the same order and compatible types, every enum
constant with the same value.  Both files are parsed independently here (no Rust toolchain in this image)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C2R = {"int": "c_int", "uint8_t": "u8", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "float": "f32", "double": "f64",
       "char": "c_char", "void": "c_void", "gsv_program": "GsvProgram", "gsv_session": "GsvSession", "gsv_ctx": "GsvCtx",
       "gsv_plan_options": "GsvPlanOptions", "gsv_program_info": "GsvProgramInfo", "gsv_session_options": "GsvSessionOptions",
       "gsv_garble_result": "GsvGarbleResult", "gsv_evaluate_io": "GsvEvaluateIo", "gsv_body_fn": "GsvBodyFn"}


def c_type_to_rust(c):
    toks = c.replace("*", " * ").split()
    base, const, i = None, False, 0
    while i < len(toks) and toks[i] != "*":
        if toks[i] == "const":
            const = True
        else:
            base = toks[i]
        i += 1
    out = C2R[base]
    while i < len(toks):
        out = ("*const " if const else "*mut ") + out
        const = False
        i += 1
        if i < len(toks) and toks[i] == "const":
            const = True
            i += 1
    return out


def parse_header():
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gsv_cuda.h")).read(), flags=re.S)
    funcs, structs, consts = {}, {}, {}
    for m in re.finditer(r"^([A-Za-z_][\w \*]*?)\b(gsv_\w+)\s*\(([^;{]*?)\)\s*;", h, flags=re.M):
        args = []
        a = " ".join(m.group(3).split())
        if a != "void":
            for x in a.split(","):
                mm = re.match(r"^(.*?)(\w+)(\[\d+\])?$", x.strip())
                t = c_type_to_rust(mm.group(1))
                args.append((mm.group(2), ("*mut " + t) if mm.group(3) else t))
        ret = m.group(1).strip()
        funcs[m.group(2)] = (args, None if ret == "void" else c_type_to_rust(ret))
    for m in re.finditer(r"typedef struct \{(.*?)\} (\w+);", h, flags=re.S):
        fields = []
        for f in m.group(1).split(";"):
            f = " ".join(f.split())
            if f:
                mm = re.match(r"^(.*?)(\w+)(\[(\d+)\])?$", f)
                t = c_type_to_rust(mm.group(1))
                fields.append((mm.group(2), f"[{t}; {mm.group(4)}]" if mm.group(4) else t))
        structs[C2R[m.group(2)]] = fields
    for m in re.finditer(r"enum \w+ \{(.*?)\};", h, flags=re.S):
        for item in m.group(1).split(","):
            if "=" in item:
                k, v = item.split("=")
                consts[k.strip()] = int(v.strip())
    return funcs, structs, consts


def parse_ffi():
    r = open(os.path.join(ROOT, "gsv-cuda", "src", "ffi.rs")).read()
    r = re.sub(r"//.*", "", r)
    funcs, structs, consts = {}, {}, {}
    ext = re.search(r'extern "C" \{(.*?)\n\}', r, flags=re.S).group(1)
    for m in re.finditer(r"pub fn (\w+)\((.*?)\)(?:\s*->\s*([^;]+))?;", ext, flags=re.S):
        args = []
        for x in filter(None, (y.strip() for y in m.group(2).split(","))):
            n, t = x.split(":", 1)
            args.append((n.strip(), " ".join(t.split())))
        funcs[m.group(1)] = (args, m.group(3).strip() if m.group(3) else None)
    for m in re.finditer(r"pub struct (\w+) \{(.*?)\}", r, flags=re.S):
        fields = []
        for x in filter(None, (y.strip() for y in m.group(2).split(",\n"))):
            x = x.rstrip(",")
            if x.startswith("pub "):
                n, t = x[4:].split(":", 1)
                fields.append((n.strip(), " ".join(t.split())))
        if fields and not fields[0][0].startswith("_"):
            structs[m.group(1)] = fields
    for m in re.finditer(r"pub const (\w+): c_int = (-?\d+);", r):
        consts[m.group(1)] = int(m.group(2))
    return funcs, structs, consts


def test_rust_ffi_matches_header():
    hf, hs, hc = parse_header()
    rf, rs, rc = parse_ffi()
    assert set(hf) == set(rf), f"functions differ: {set(hf) ^ set(rf)}"
    for name in hf:
        assert hf[name] == rf[name], f"{name}: header {hf[name]} vs ffi.rs {rf[name]}"
    assert set(hs) == set(rs), f"structs differ: {set(hs) ^ set(rs)}"
    for name in hs:
        assert hs[name] == rs[name], f"struct {name}: header {hs[name]} vs ffi.rs {rs[name]}"
    assert hc == rc


def test_ctypes_structs_match_header(gsv):
    """The Python binding's ctypes structures have the header's fields, in order."""
    _, hs, _ = parse_header()
    pairs = {"GsvPlanOptions": gsv._PlanOptions, "GsvProgramInfo": gsv._ProgramInfo, "GsvSessionOptions": gsv._SessionOptions,
             "GsvGarbleResult": gsv._GarbleResult, "GsvEvaluateIo": gsv._EvaluateIO}
    for name, cls in pairs.items():
        assert [f[0] for f in hs[name]] == [f[0] for f in cls._fields_], name


def test_library_exports_every_header_symbol(gsv):
    hf, _, _ = parse_header()
    lib = gsv.load_library()
    for name in hf:
        assert hasattr(lib, name), f"libgsv_cuda.so does not export {name}"
