#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle/_ref): rewrites the reference's GLSL 3.30 hot-path shaders into C++ that compiles against
glsl_types.hpp, so that the reference's own shader text runs on the CPU and pins the hand-written oracle (oracle/*.c).

The shaders are read where they lie (/root/reference/src/shader); the generated C++ goes to oracle/_ref/gen/ (git-ignored).
Nothing of the reference is copied into the repository.

The rewrite is mechanical and keeps every expression as written:
  * #version / #pragma dropped, #include "shader/x.glsl" inlined, comments removed
  * layout(...) / flat / uniform / in / out qualifiers dropped at file scope; globals become members of one struct per
    stage (`struct Shader : glsl::StageBase`), functions become its methods; `out` names are recorded so that
    EmitVertex() can snapshot them (geometry stages)
  * interface blocks `out SURFEL {...} vs_out;` / `in SURFEL {...} gs_in[];` become a nested struct + a member (array of 1)
  * floating literals get an `f` suffix (GLSL literals are fp32; C++ ones would silently promote expressions to double)
  * `discard;` sets a flag and returns

usage: glsl2cpp.py <reference_src_dir> <out_dir> shader/a.vert shader/b.geom ...
"""
import hashlib
import os
import re
import sys

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)[fF]?(?![\w.])")


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def load(src_dir, rel, seen=None):
    seen = seen or []
    text = open(os.path.join(src_dir, rel)).read()
    digest = [(rel, hashlib.sha256(text.encode()).hexdigest()[:16])]
    out = []
    for line in strip_comments(text).splitlines():
        m = re.match(r'\s*#\s*include\s+"([^"]+)"', line)
        if m:
            if m.group(1) not in seen:
                seen.append(m.group(1))
                sub, d = load(src_dir, m.group(1), seen)
                out.append(sub)
                digest += d
            continue
        if re.match(r"\s*#\s*(version|pragma)", line):
            continue
        out.append(line)
    return "\n".join(out), digest


def chunks(text):
    """file-scope pieces: declarations (ending in ';' at depth 0) and function definitions (ending in their '}')"""
    depth, start, i, n = 0, 0, 0, len(text)
    while i < n:
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                head = text[start:text.index("{", start)]
                if "(" in head:                       # a function body just closed
                    yield text[start:i + 1].strip()
                    start = i + 1
        elif c == ";" and depth == 0:
            piece = text[start:i + 1].strip()
            if piece != ";":
                yield piece
            start = i + 1
        i += 1
    rest = text[start:].strip()
    if rest:
        raise SystemExit("glsl2cpp: trailing text not understood: %r" % rest[:80])


def rewrite_decl(piece, outs, refl):
    """one file-scope declaration -> member declaration text; `refl` collects what the generic software GL
    (ref_harness/full) needs to know about the stage's interface: uniforms, inputs, outputs, locations"""
    loc = re.search(r"layout\s*\(\s*location\s*=\s*(\d+)\s*\)", piece)
    loc = int(loc.group(1)) if loc else -1
    p = re.sub(r"layout\s*\([^)]*\)\s*", "", piece).strip()
    if re.fullmatch(r"(in|out)\s*;", p):
        m = re.search(r"layout\s*\(\s*(points|line_strip|triangle_strip)\b", piece)
        if m and re.fullmatch(r"out\s*;", p):
            refl.append(dict(kind="out_primitive", prim=m.group(1)))
        return None                                    # layout(points) in;  /  layout(..., max_vertices = N) out;
    p = re.sub(r"^\s*flat\s+", "", p)
    m = re.fullmatch(r"(in|out)\s+(\w+)\s*\{(.*)\}\s*(\w+)\s*(\[\s*\])?\s*;", p, flags=re.S)
    if m:
        qual, block, body, inst, arr = m.groups()
        if qual == "out":
            outs.append((block, inst))
        refl.append(dict(kind="block", qual=qual, typ=block, name=inst))
        return "struct %s {%s};\n  %s %s%s;" % (block, body, block, inst, "[1]" if arr else "")
    m = re.fullmatch(r"(in|out)\s+(\w+)\s+(\w+)\s*;", p)
    if m:
        qual, typ, name = m.groups()
        if qual == "out":
            outs.append((typ, name))
        refl.append(dict(kind="var", qual=qual, typ=typ, name=name, loc=loc))
        return "%s %s%s;" % (typ, name, " = %s()" % typ if typ in ("int", "float", "bool") else "")
    m = re.fullmatch(r"uniform\s+(\w+)\s+(\w+)\s*;", p)
    if m:
        typ, name = m.groups()
        refl.append(dict(kind="uniform", typ=typ, name=name))
        return "%s %s%s;" % (typ, name, " = %s()" % typ if typ in ("int", "float", "bool") else "")
    return p                                           # const / plain global with or without initialiser


def transpile(src_dir, rel):
    text, digest = load(src_dir, rel)
    text = FLOAT_LIT.sub(lambda m: m.group(1) + "f", text)
    text = re.sub(r"\bdiscard\s*;", "{ discarded_ = true; return; }", text)
    outs, body, refl = [], [], []
    for piece in chunks(text):
        head = piece.split("{", 1)[0]
        if "(" in head and piece.endswith("}"):
            body.append(piece)                         # function -> method, text unchanged
        else:
            d = rewrite_decl(piece, outs, refl)
            if d:
                body.append(d)
    name = "ref_" + os.path.basename(rel).replace(".", "_")
    snap = "".join(" e.%s = %s;" % (n, n) for _, n in outs)
    fields = "".join(" %s %s;" % (t, n) for t, n in outs)
    lines = ["// GENERATED by oracle/ref_harness/glsl2cpp.py -- derived from the reference's source, never committed.",
             "// sources: " + ", ".join("%s sha256:%s" % d for d in digest),
             "#pragma once", '#include "glsl_types.hpp"', "namespace glsl { namespace %s {" % name,
             "struct Shader : StageBase {"]
    lines += ["  " + b for b in body]
    lines += ["  struct Emitted { vec4 gl_Position;%s };" % fields,
              "  std::vector<Emitted> emitted_; std::vector<int> prim_end_;",
              "  void EmitVertex() { Emitted e; e.gl_Position = gl_Position;%s emitted_.push_back(e); }" % snap,
              "  void EndPrimitive() { prim_end_.push_back((int)emitted_.size()); }",
              "};"]
    lines += reflection(rel, refl, outs)
    lines += ["}}  // namespace", ""]
    return name, "\n".join(lines), digest


NCOMP = {"float": 1, "vec2": 2, "vec3": 3, "vec4": 4}
SAMPLERS = {"sampler2DRect": "rect", "samplerBuffer": "buffer"}


def reflection(rel, refl, outs):
    """type-erased adapter (sgl::AnyStage) of one transpiled stage, compiled only into the generic software GL
    (-DSGL_REFLECT): uniforms by name, vertex attributes / fragment outputs by location, varyings by name"""
    L = ["#ifdef SGL_REFLECT", "struct Any : sgl::AnyStage {", "  Shader s;"]
    samplers = [r for r in refl if r["kind"] == "uniform" and r["typ"] in SAMPLERS]
    L.append("  int unit_[%d] = {%s};" % (max(1, len(samplers)), ", ".join("0" for _ in range(max(1, len(samplers))))))
    L.append("  glsl::StageBase& base() override { return s; }")
    L.append("  void run() override { s.main(); }")
    L.append("  bool set_uniform(const std::string& n, const sgl::UVal& v) override {")
    for k, r in enumerate(samplers):
        L.append('    if (n == "%s") { unit_[%d] = v.as_int(); return true; }' % (r["name"], k))
    for r in refl:
        if r["kind"] == "uniform" and r["typ"] not in SAMPLERS:
            L.append('    if (n == "%s") { sgl::assign_uniform(s.%s, v); return true; }' % (r["name"], r["name"]))
    L.append("    return false;")
    L.append("  }")
    L.append("  void bind_samplers(const sgl::Units& u) override {")
    for k, r in enumerate(samplers):
        L.append("    s.%s = u.%s(unit_[%d]);" % (r["name"], SAMPLERS[r["typ"]], k))
    L.append("  }")
    L.append("  void set_attribute(int loc, const sgl::AttrVal& a) override {")
    for r in refl:
        if r["kind"] == "var" and r["qual"] == "in" and r["loc"] >= 0:
            L.append("    if (loc == %d) sgl::load_attr(s.%s, a);" % (r["loc"], r["name"]))
    L.append("  }")
    for qual, fn in (("in", "ins"), ("out", "outs")):
        L.append("  std::vector<sgl::Var> %s() override {" % fn)
        L.append("    std::vector<sgl::Var> v;")
        for r in refl:
            if r.get("qual") != qual:
                continue
            if r["kind"] == "block":   # matched between stages by BLOCK name; an input block is an array of one
                L.append('    v.push_back(sgl::Var{"%s", (void*)sgl::first_elem(s.%s), sizeof(Shader::%s), 0, -1});' %
                         (r["typ"], r["name"], r["typ"]))
            else:
                L.append('    v.push_back(sgl::Var{"%s", (void*)&s.%s, sizeof(s.%s), %d, %d});' %
                         (r["name"], r["name"], r["name"], NCOMP.get(r["typ"], 0), r["loc"]))
        L.append("    return v;")
        L.append("  }")
    # emitted vertices of a geometry stage (also used for a vertex stage's outputs snapshot: none needed there)
    prim = [r["prim"] for r in refl if r["kind"] == "out_primitive"]
    L.append("  bool emits_triangle_strips() override { return %s; }" % ("true" if prim == ["triangle_strip"] else "false"))
    L.append("  int n_emitted() override { return (int)s.emitted_.size(); }")
    L.append("  void clear_emitted() override { s.emitted_.clear(); s.prim_end_.clear(); }")
    L.append("  const std::vector<int>& prim_ends() override { return s.prim_end_; }")
    L.append("  const glsl::vec4& emitted_position(int i) override { return s.emitted_[i].gl_Position; }")
    L.append("  const char* emitted_data(int i) override { return (const char*)&s.emitted_[i]; }")
    L.append("  std::vector<sgl::Var> emitted_vars() override {")
    L.append("    std::vector<sgl::Var> v;")
    for typ, name in outs:
        L.append('    v.push_back(sgl::Var{"%s", (void*)offsetof(Shader::Emitted, %s), sizeof(((Shader::Emitted*)0)->%s), %d, -1});' %
                 (name if typ in NCOMP or typ in ("int", "bool", "uint") else typ, name, name, NCOMP.get(typ, 0)))
    L.append("    return v;")
    L.append("  }")
    L += ["};", 'static sgl::Registrar reg_("%s", [] { return (sgl::AnyStage*)new Any(); });' % rel, "#endif"]
    return L


def swizzle_inc(n):
    names = ["xyzw"[:n], "rgba"[:n]]
    out = []
    for comps in names:
        for k in (2, 3, 4):
            idx = [0] * k
            while True:
                nm = "".join(comps[i] for i in idx)
                out.append("Swz<%d, %s> %s;" % (n, ", ".join(map(str, idx)), nm))
                j = k - 1
                while j >= 0:
                    idx[j] += 1
                    if idx[j] < n:
                        break
                    idx[j] = 0
                    j -= 1
                if j < 0:
                    break
    return "\n".join(out) + "\n"


def main():
    src_dir, out_dir, rels = sys.argv[1], sys.argv[2], sys.argv[3:]
    os.makedirs(out_dir, exist_ok=True)
    for n in (2, 3, 4):
        open(os.path.join(out_dir, "swizzles%d.inc" % n), "w").write(swizzle_inc(n))
    manifest = []
    for rel in rels:
        name, cpp, digest = transpile(src_dir, rel)
        open(os.path.join(out_dir, name + ".hpp"), "w").write(cpp)
        manifest.append("%s <- %s" % (name, ", ".join("%s@%s" % d for d in digest)))
    open(os.path.join(out_dir, "all_stages.inc"), "w").write(
        "".join('#include "ref_%s.hpp"\n' % os.path.basename(r).replace(".", "_") for r in rels))
    open(os.path.join(out_dir, "MANIFEST.txt"), "w").write("\n".join(manifest) + "\n")


if __name__ == "__main__":
    main()
