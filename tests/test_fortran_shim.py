"""CPU: static consistency of fortran/fcp_b200.f90 (the ISO_C_BINDING shim a maintainer of the reference links) with
include/fcp.h.  No Fortran compiler exists in this image, so the shim cannot be compiled here; what CAN be checked without one is the part
a compiler would NOT catch anyway -- drift between the two descriptions of the ABI:

  * every `bind(c, name='...')` interface names a function the header declares, with the same number of arguments and, position by
    position, the same passing class (int / int32 / int64 / double / float by value, pointer otherwise) and the same return class;
  * every `type, bind(c)` has the header struct's fields in the same order with the same types (arrays included);
  * the field-id enumerators and the integer constants have the header's names and values;
and the part a compiler WOULD catch first: balanced program units, `implicit none` in every module, and no name used in an executable
statement of a module procedure that is neither declared locally, nor at module level, nor reachable through a `use`."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F90 = os.path.join(ROOT, "fortran", "fcp_b200.f90")
HDR = os.path.join(ROOT, "include", "fcp.h")


# ----------------------------------------------------------------------------------------------------------------- C side
def _c_text():
    s = open(HDR).read()
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    s = re.sub(r"//[^\n]*", " ", s)
    return s


def _c_class(decl: str) -> str:
    decl = decl.strip()
    if "*" in decl or "[" in decl:
        return "ptr"
    t = re.sub(r"\bconst\b", "", decl).split()
    base = " ".join(t[:-1]) if len(t) > 1 else t[0]
    return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "double": "f64", "float": "f32", "unsigned int": "i32"}[base]


def c_functions():
    out = {}
    for m in re.finditer(r"\b(int|const char \*|void)\s*(fcp_\w+)\s*\(([^)]*)\)\s*;", _c_text()):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        cls = [] if args in ("void", "") else [_c_class(a) for a in args.split(",")]
        out[name] = ("ptr" if "*" in ret else "i32", cls)
    return out


def c_structs():
    out = {}
    for m in re.finditer(r"typedef struct\s*\{(.*?)\}\s*(fcp_\w+)\s*;", _c_text(), flags=re.S):
        fields = []
        for stmt in m.group(1).split(";"):
            stmt = " ".join(stmt.split())
            if not stmt:
                continue
            mm = re.match(r"(const\s+)?(\w+(?:\s+\w+)?)\s+(.*)", stmt)
            base = mm.group(2)
            for d in mm.group(3).split(","):
                d = d.strip()
                arr = re.search(r"\[(\d+)\]", d)
                name = re.sub(r"[\*\s]|\[\d+\]", "", d)
                kind = "ptr" if "*" in d else _c_class(base + " x")
                fields.append((name, kind, int(arr.group(1)) if arr else 0))
        out[m.group(2)] = fields
    return out


def c_constants():
    out = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(FCP_\w+)\s+(-?\d+)\b", _c_text())}
    for m in re.finditer(r"enum\s*\{(.*?)\}", _c_text(), flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = item.split("=")
                nxt = int(v.strip(), 0)
                out[k.strip()] = nxt
            else:
                out[item] = nxt
            nxt += 1
    return out


# ------------------------------------------------------------------------------------------------------------- Fortran side
def f_lines():
    """logical lines: comments stripped, continuations joined, lower-cased"""
    out, cur = [], ""
    for raw in open(F90).read().splitlines():
        line, q = "", None
        for ch in raw:                       # strip a trailing comment, minding character literals
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "!":
                break
            line += ch
        line = line.strip()
        if not line:
            continue
        if line.startswith("&"):
            line = line[1:].lstrip()
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        full, q, stmt = (cur + line).strip(), None, ""
        cur = ""
        for ch in full:                      # several statements on one line
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == ";":
                if stmt.strip():
                    out.append(stmt.strip())
                stmt = ""
                continue
            stmt += ch
        if stmt.strip():
            out.append(stmt.strip())
    return out


F_KIND = {"c_int": "i32", "c_int32_t": "i32", "c_int64_t": "i64", "c_double": "f64", "c_float": "f32"}


def _f_decl(line: str):
    """'integer(c_int), value :: a, b(*)' -> [(name, class, arraylen)]"""
    m = re.match(r"(integer|real|type|character)\s*\(([^)]*)\)\s*(.*?)::\s*(.*)", line, flags=re.I)
    if not m:
        return []
    base, kind, attrs, names = m.group(1).lower(), m.group(2).strip(), m.group(3).lower(), m.group(4)
    res = []
    for nm in re.split(r",(?![^()]*\))", names):
        nm = nm.strip()
        arr = re.search(r"\(([^)]*)\)", nm)
        name = re.sub(r"\(.*\)", "", nm).strip()
        if base == "type":
            cls = "ptr"                                   # type(c_ptr) by value IS the pointer; a derived type by reference is a pointer to it
        elif base == "character":
            cls = "ptr"
        elif arr or "value" not in attrs:
            cls = "ptr"
        else:
            cls = F_KIND[kind.replace("kind=", "").strip()]
        n = int(arr.group(1)) if arr and arr.group(1).isdigit() else 0
        res.append((name, cls, n, F_KIND.get(kind, "ptr") if base != "type" else "ptr"))
    return res


def f_interfaces(lines):
    out, i = {}, 0
    while i < len(lines):
        m = re.match(r"(function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\s*\(\s*c\s*,\s*name\s*=\s*'(\w+)'\s*\)(?:\s*result\s*\((\w+)\))?", lines[i], flags=re.I)
        if not m:
            i += 1
            continue
        args = [a.strip() for a in m.group(3).split(",") if a.strip()]
        res, decl = m.group(5), {}
        i += 1
        while not re.match(r"end\s*(function|subroutine)", lines[i], flags=re.I):
            for name, cls, _, _ in _f_decl(lines[i]):
                decl[name.lower()] = cls if name.lower() != (res or "").lower() else ("ptr" if lines[i].lower().startswith("type") else F_KIND[re.search(r"\((\w+)\)", lines[i]).group(1)])
            i += 1
        out[m.group(4)] = (decl.get((res or "").lower(), "void"), [decl[a.lower()] for a in args], m.group(2))
    return out


def f_types(lines):
    out, i = {}, 0
    while i < len(lines):
        m = re.match(r"type\s*,\s*bind\s*\(\s*c\s*\)\s*::\s*(\w+)", lines[i], flags=re.I)
        if not m:
            i += 1
            continue
        fields = []
        i += 1
        while not re.match(r"end\s*type", lines[i], flags=re.I):
            for name, _, n, elem in _f_decl(lines[i]):
                fields.append((name, elem, n))
            i += 1
        out[m.group(1)] = fields
    return out


# ----------------------------------------------------------------------------------------------------------------------- tests
def test_every_interface_matches_the_header():
    cf, ff = c_functions(), f_interfaces(f_lines())
    assert len(ff) >= 35, sorted(ff)
    for name, (fret, fargs, fname) in ff.items():
        assert name in cf, f"{name}: bound in the shim but not declared in include/fcp.h"
        assert fname == name, f"{fname}: Fortran name differs from its binding label {name}"
        cret, cargs = cf[name]
        assert fret == cret, f"{name}: return class {fret} vs {cret}"
        assert len(fargs) == len(cargs), f"{name}: {len(fargs)} arguments in the shim, {len(cargs)} in the header"
        for k, (a, b) in enumerate(zip(fargs, cargs)):
            assert a == b, f"{name}: argument {k + 1} is {a} in the shim and {b} in the header"


def test_every_bind_c_type_matches_the_header():
    cs, fs = c_structs(), f_types(f_lines())
    assert set(fs) <= set(cs), set(fs) - set(cs)
    assert {"fcp_mesh_desc", "fcp_report", "fcp_simple_params", "fcp_piso_params", "fcp_uvw_params", "fcp_scalar_params"} <= set(fs)
    for name, ffields in fs.items():
        cfields = cs[name]
        assert [f[0].lower() for f in ffields] == [c[0].lower() for c in cfields], f"{name}: field order\n{[f[0] for f in ffields]}\n{[c[0] for c in cfields]}"
        for (fn, fk, fa), (cn, ck, ca) in zip(ffields, cfields):
            assert (fk, fa) == (ck, ca), f"{name}.{cn}: {fk}[{fa}] in the shim, {ck}[{ca}] in the header"


def test_constants_match_the_header():
    cc = c_constants()
    text = " ".join(f_lines())
    seen = 0
    for m in re.finditer(r"\b(FCP_[A-Z0-9_]+)\s*=\s*(-?\d+)", text):
        assert m.group(1) in cc and cc[m.group(1)] == int(m.group(2)), f"{m.group(1)} = {m.group(2)} in the shim, {cc.get(m.group(1))} in the header"
        seen += 1
    assert seen >= 25
    en = re.search(r"enum\s*,\s*bind\s*\(c\)\s*enumerator\s*::(.*?)end enum", text, flags=re.I | re.S).group(1)
    names = [re.sub(r"=.*", "", t).strip() for t in en.split(",")]
    names = [re.sub(r"^enumerator\s*::\s*", "", n, flags=re.I) for n in names]
    for k, nme in enumerate(names):
        assert cc.get(nme) == k, f"field id {nme}: position {k} in the shim, {cc.get(nme)} in the header"
    import fcb200  # noqa: F401
    from fcb200 import lib as L
    assert [n[len("FCP_F_"):] for n in names] == L.FIELDS



REF = "/root/reference/src"


def _ref_module_file(mod: str):
    for dp_, _, fs in os.walk(REF):
        for fn in fs:
            if fn.lower().endswith((".f90", ".f95", ".f")):
                p = os.path.join(dp_, fn)
                try:
                    txt = open(p, errors="replace").read()
                except OSError:
                    continue
                if re.search(r"^\s*module\s+%s\s*(!.*)?$" % re.escape(mod), txt, flags=re.I | re.M):
                    return p
    return None


def ref_module_names(mod: str, seen=None):
    """names a `use mod` makes visible: the module's own declarations and procedures plus what it re-exports from the modules it uses"""
    seen = set() if seen is None else seen
    if mod.lower() in seen or mod.lower() in ("iso_c_binding", "mpi", "omp_lib", "ieee_arithmetic"):
        return set()
    seen.add(mod.lower())
    p = _ref_module_file(mod)
    assert p, f"module {mod} not found in the reference"
    names, inside, after_contains, depth = set(), False, False, 0
    cur = ""
    for raw in open(p, errors="replace").read().splitlines():
        line = raw.split("!")[0].strip()
        if not line:
            continue
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        line = (cur + line.lstrip("&")).strip()
        cur = ""
        low = line.lower()
        if not inside:
            inside = bool(re.match(r"module\s+%s\s*$" % re.escape(mod.lower()), low))
            continue
        if re.match(r"end\s*module", low):
            break
        if low == "contains" and depth == 0:
            after_contains = True
            continue
        if after_contains:
            m = re.match(r"(?:(?:pure|elemental|recursive|integer|real\s*\(\w+\)|logical)\s+)*(subroutine|function)\s+(\w+)", low)
            if m and depth == 0:
                names.add(m.group(2))
            if m:
                depth += 1
            elif re.match(r"end\s*(subroutine|function)", low):
                depth -= 1
            continue
        m = re.match(r"use\s+(\w+)\s*(?:,\s*only\s*:\s*(.*))?$", low)
        if m:
            if m.group(2):
                names |= {t.split("=>")[0].strip() for t in m.group(2).split(",")}
            else:
                names |= ref_module_names(m.group(1), seen)
            continue
        m = re.match(r"interface\s+(\w+)", low)
        if m:
            names.add(m.group(1))
        m = re.match(r"type\s*(?:,[^:]*)?(?:::)?\s*(\w+)\s*$", low)
        if m and not low.startswith("type("):
            names.add(m.group(1))
        if "::" in low:
            for t in re.split(r",(?![^()]*\))", low.split("::", 1)[1]):
                names.add(re.sub(r"[\(=].*", "", t).strip())
    return names


# external (non-module) subroutines of the reference the shim calls through an implicit interface, with the file that defines each
EXTERNAL_SUBROUTINES = {"continuityerrors": "finiteVolume/fvEqnDiscretization/Pressure/continuityErrors.f90",
                        "constant_mass_flow_forcing": "cappuccino/constant_mass_flow_forcing.f90"}

FORTRAN_WORDS = set("""if then else elseif endif end do enddo while select case default call return stop exit cycle subroutine function module contains use
implicit none only intent in out inout integer real character logical type class parameter allocatable dimension target pointer save value
optional result interface procedure public private allocate deallocate allocated associated present size trim adjustl len len_trim abs max
min sqrt sum maxval minval real int dble nint mod sign print write read open close format kind true false and or not eq ne lt le gt ge eqv
neqv c_loc c_f_pointer c_associated c_null_ptr c_null_char c_ptr c_int c_int32_t c_int64_t c_double c_float c_char c_funloc iso_c_binding
intrinsic import continue go to where elsewhere forall stat errmsg unit fmt advance iostat exp log tanh achar char index present huge tiny
epsilon merge reshape shape lbound ubound any all count dot_product matmul transpose null nullify block data enum enumerator bind c name
error dp""".split())


def test_module_procedures_use_only_declared_names():
    """A poor man's `implicit none` check for the executable part (module fcp_backend): every identifier in a procedure body must be a local
    declaration, a dummy argument, a module-level entity of this file, or come from a `use` of one of the REFERENCE's modules (listed by name
    in the `use ..., only:` clauses, which is how the shim imports the reference's arrays)."""
    lines = f_lines()
    start = next(i for i, l in enumerate(lines) if re.match(r"module\s+fcp_backend", l, flags=re.I))
    body = lines[start:]
    assert any(re.match(r"implicit\s+none", l, flags=re.I) for l in body[:40])
    assert re.match(r"implicit\s+none", next(l for l in lines[:10] if l.lower().startswith("implicit")), flags=re.I)
    # names visible module-wide: everything declared or imported before `contains`, every procedure name, everything module fcp_b200 exports
    glob = set()
    decl_re = re.compile(r"(integer|real|logical|character|type\s*\()[^:]*::\s*(.*)", flags=re.I)

    def declared(line):
        m = decl_re.match(line)
        if not m:
            return []
        return [re.sub(r"[\(=].*", "", t).strip().lower() for t in re.split(r",(?![^()]*\))", m.group(2))]
    for l in lines:
        for nm in declared(l) if lines.index(l) < start else []:
            glob.add(nm)
        m = re.match(r"(?:(?:integer|real|logical)\s*(?:\([^)]*\))?\s+)?(?:function|subroutine)\s+(\w+)", l, flags=re.I)
        if m:
            glob.add(m.group(1).lower())
        m = re.match(r"type\s*,\s*bind\s*\(\s*c\s*\)\s*::\s*(\w+)", l, flags=re.I)
        if m:
            glob.add(m.group(1).lower())
    glob |= {k.lower() for k in c_constants()}
    ci = next(i for i, l in enumerate(body) if l.lower() == "contains")
    import pytest
    if not os.path.isdir(REF):
        pytest.skip("the reference tree (for the names its modules export) is not on this machine")
    problems = []

    def check_only(where, mod, items):
        vis = set()
        exported = ref_module_names(mod) if mod.lower() not in ("iso_c_binding", "fcp_b200") else None
        for t in items.split(","):
            parts = [x.strip().lower() for x in t.split("=>")]
            vis.add(parts[0])
            if exported is not None and parts[-1] not in exported:
                problems.append(f"{where}: `use {mod}, only: {parts[-1]}` -- the reference's module {mod} has no such entity")
        return vis
    for l in body[:ci]:
        glob |= set(declared(l))
        m = re.match(r"use\s*(?:,\s*intrinsic\s*::)?\s*(\w+)\s*(?:,\s*only\s*:\s*(.*))?$", l, flags=re.I)
        if m and m.group(2):
            glob |= check_only("module fcp_backend", m.group(1), m.group(2))
        elif m and m.group(1).lower() not in ("iso_c_binding", "fcp_b200"):
            glob |= ref_module_names(m.group(1))
    # walk the procedures
    i, nproc = ci + 1, 0
    while i < len(body):
        m = re.match(r"subroutine\s+(\w+)\s*(?:\(([^)]*)\))?", body[i], flags=re.I)
        if not m:
            i += 1
            continue
        nproc += 1
        pname = m.group(1)
        local = {a.strip().lower() for a in (m.group(2) or "").split(",") if a.strip()}
        j = i + 1
        while not re.match(r"end\s*subroutine", body[j], flags=re.I):
            l = body[j]
            mu = re.match(r"use\s+\w+\s*(?:,\s*only\s*:\s*(.*))?$", l, flags=re.I)
            if mu:
                if mu.group(1):
                    local |= check_only(pname, re.match(r"use\s+(\w+)", l, flags=re.I).group(1), mu.group(1))
                else:
                    problems.append(f"{pname}: `{l}` imports a whole module (cannot be checked): use an only-list")
            elif declared(l):
                local |= set(declared(l))
                rhs = l.split("::", 1)[1]
                for tok in re.findall(r"[a-z_]\w*", re.sub(r"'[^']*'|\"[^\"]*\"", " ", rhs.lower())):
                    if tok not in local and tok not in glob and tok not in FORTRAN_WORDS:
                        problems.append(f"{pname}: `{tok}` in a declaration is not visible")
            elif not re.match(r"implicit\s+none", l, flags=re.I):
                code = re.sub(r"'[^']*'|\"[^\"]*\"", " ", l.lower())
                code = re.sub(r"(?<![\w.])\d+\.?\d*(?:[de][+-]?\d+)?(?:_\w+)?", " ", code)       # numeric literals with kind suffixes
                code = re.sub(r"%\s*\w+", " ", code)                                    # derived-type components
                code = re.sub(r"\bcall\s+(\w+)", lambda mm: " " if (mm.group(1) in glob or mm.group(1) in EXTERNAL_SUBROUTINES) else mm.group(0), code)
                code = re.sub(r"\.(?:and|or|not|eq|ne|lt|le|gt|ge|true|false|eqv|neqv)\.", " ", code)
                for tok in re.findall(r"[a-z_]\w*", code):
                    if tok not in local and tok not in glob and tok not in FORTRAN_WORDS:
                        problems.append(f"{pname}: `{tok}` is not declared, imported or a module entity   [{l[:90]}]")
            j += 1
        i = j + 1
    assert nproc >= 20
    for name, rel in EXTERNAL_SUBROUTINES.items():
        txt = open(os.path.join(REF, rel), errors="replace").read().lower()
        assert re.search(r"^\s*subroutine\s+%s\b" % name, txt, flags=re.M) and not re.search(r"^\s*module\s+\w+\s*$", txt, flags=re.M), name
    assert not problems, "\n".join(sorted(set(problems)))


def test_program_units_are_balanced():
    stack = []
    for l in f_lines():
        low = l.lower()
        mend = re.match(r"end\s*(module|subroutine|function|interface|type|enum)\b", low)
        if mend:
            assert stack and stack[-1] == mend.group(1), f"`{l}` closes {stack[-1] if stack else 'nothing'}"
            stack.pop()
        elif re.match(r"module\s+(?!procedure)\w+\s*$", low):
            stack.append("module")
        elif re.match(r"(?:(?:integer|real|logical|pure|elemental|recursive)\s*(?:\([^)]*\))?\s+)*(subroutine|function)\s+\w+", low):
            stack.append(re.search(r"(subroutine|function)", low).group(1))
        elif re.match(r"interface\b", low):
            stack.append("interface")
        elif re.match(r"type\s*(,|::|\s+\w+\s*$)", low) and not low.startswith("type("):
            stack.append("type")
        elif re.match(r"enum\s*,", low):
            stack.append("enum")
    assert not stack, stack
    # if / do / select nesting inside the procedures
    depth = {"if": 0, "do": 0, "select": 0}
    for l in f_lines():
        low = l.lower()
        if re.match(r"(\w+\s*:\s*)?if\s*\(.*\)\s*then$", low):
            depth["if"] += 1
        elif re.match(r"end\s*if\b", low):
            depth["if"] -= 1
        elif re.match(r"(\w+\s*:\s*)?do\b(?!uble)", low):
            depth["do"] += 1
        elif re.match(r"end\s*do\b", low):
            depth["do"] -= 1
        elif re.match(r"select\s+case", low):
            depth["select"] += 1
        elif re.match(r"end\s*select", low):
            depth["select"] -= 1
        assert min(depth.values()) >= 0, l
    assert depth == {"if": 0, "do": 0, "select": 0}, depth
