"""The ctypes prototypes in _native.py must agree with the C prototypes in include/tb200.h argument by argument
(a mismatch is silent undefined behaviour at call time), and the ctypes mirror structs with the C structs field by
field. Pure host-side check: parses the header, needs no GPU."""
import ctypes as C
import os
import re

from pytorch_toolbelt_b200 import _native as N

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(REPO, "include", "tb200.h")).read()


def _strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def _kind_of_c(decl: str) -> str:
    decl = decl.strip()
    if "*" in decl:
        return "ptr"
    base = re.sub(r"\b(const|unsigned|signed)\b", "", decl).split()
    ctype = base[0] if base else ""
    return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "float": "f32", "double": "f64", "size_t": "u64",  # c_size_t is c_uint64 on LP64
           
            "uint64_t": "u64"}.get(ctype, "?" + ctype)


def _kind_of_ctypes(t) -> str:
    if t is C.c_void_p or t is C.c_char_p or (isinstance(t, type) and issubclass(t, C._Pointer)):
        return "ptr"
    return {C.c_int: "i32", C.c_int64: "i64", C.c_float: "f32", C.c_double: "f64", C.c_size_t: "u64", C.c_uint64: "u64"}.get(t, f"?{t}")


def _prototypes():
    text = _strip_comments(HEADER)
    for m in re.finditer(r"TB_API\s+([\w\s\*]+?)\b(tb_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        args = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        yield name, ret, args


def test_every_prototype_matches_its_ctypes_signature():
    protos = list(_prototypes())
    assert len(protos) >= 24 and {p[0] for p in protos} == set(N.SIGNATURES)
    problems = []
    for name, ret, args in protos:
        restype, argtypes = N.SIGNATURES[name]
        want = [_kind_of_c(a) for a in args]
        got = [_kind_of_ctypes(t) for t in argtypes]
        if want != got:
            problems.append(f"{name}: header {want} vs ctypes {got}")
        ret_kind = "ptr" if "*" in ret else _kind_of_c(ret)
        if ret_kind != _kind_of_ctypes(restype):
            problems.append(f"{name}: return {ret_kind} vs {_kind_of_ctypes(restype)}")
    assert not problems, "\n".join(problems)


def _c_struct_fields(name):
    text = _strip_comments(HEADER)
    m = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), text, flags=re.S)
    assert m, f"struct {name} not found in the header"
    fields = []
    for stmt in m.group(1).split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        if "*" in stmt:
            fields.append((re.findall(r"(\w+)\s*$", stmt)[0], "ptr"))
            continue
        ctype = " ".join(stmt.split()[:-1]) if "," not in stmt else stmt.split()[0]
        for ident in stmt[len(stmt.split()[0]):].split(","):
            fields.append((ident.strip(), _kind_of_c(ctype)))
    return fields


def test_mirror_structs_match_the_header():
    for cname, mirror in [("tb_weight", N.Weight), ("tb_grid", N.Grid), ("tb_focal_params", N.FocalParams),
                          ("tb_dice_params", N.DiceParams), ("tb_merge_extra", N.MergeExtra), ("tb_peer_plan", N.PeerPlan), ("tb_elem_loss_params", N.ElemLossParams)]:
        want = _c_struct_fields(cname)
        got = [(f[0], _kind_of_ctypes(f[1])) for f in mirror._fields_]
        assert want == got, f"{cname}: header {want} vs ctypes {got}"
