"""TEST INFRASTRUCTURE: builds tests/cusim/_build/libsuma_b200_sim.so -- the CUDA sources of libsuma_b200 compiled for the
CPU executor of cusim.hpp (see there). The sources are taken as they are; two mechanical rewrites make them C++:

  kernel<<<grid, block, smem, stream>>>(args);   ->  cusim::launch("kernel", cusim::cfg(grid, block, smem, stream), [&] { kernel(args); });
  asm volatile("mov.u64 %0, %%globaltimer;" ...) ->  the CPU clock; any other inline PTX (the TMA staging of
                                                     k_preprocess_tile<true>) -> a call that aborts: the executor reports no
                                                     TMA unit, so the library takes its plain-load instantiation
and the one cudaLaunchCooperativeKernel call site becomes two direct cooperative launches. Every rewrite asserts that it
found what it expects, so a change of the sources fails the build instead of silently testing something else.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "semantic_suma_b200", "csrc")
# CUSIM_ASAN=1: AddressSanitizer build -- every "device" buffer is a heap block, so an out-of-bounds kernel access is
# reported with its source line (a memcheck of the kernels); run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
ASAN = os.environ.get("CUSIM_ASAN") == "1"
# CUSIM_TSAN=1: ThreadSanitizer build = a racecheck of the kernels (CUDA threads are TSan fibers, see cusim_rt.cpp);
# run python with LD_PRELOAD=$(gcc -print-file-name=libtsan.so)
TSAN = os.environ.get("CUSIM_TSAN") == "1"
# CUSIM_COV=1: gcov instrumentation -- which lines of csrc/*.cu the tests reach (scratch/cusim_coverage.sh)
COV = os.environ.get("CUSIM_COV") == "1"
BUILD = os.path.join(HERE, "_build", "asan" if ASAN else ("tsan" if TSAN else ("cov" if COV else "plain")))
GEN = os.path.join(BUILD, "gen")
OUT = os.path.join(BUILD, "libsuma_b200_sim.so")
SOURCES = ["sb_preprocess.cu", "sb_icp.cu", "sb_map.cu", "sb_api.cu"]
CXXFLAGS = ["-std=c++17", "-O2", "-g", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-fno-strict-aliasing",
            "-mtls-dialect=gnu2", "-Wno-unknown-pragmas", "-Wno-attributes", "-D__CUDA_ARCH__=1000", "-DCUSIM=1"]
if ASAN:
    CXXFLAGS += ["-fsanitize=address", "-fno-omit-frame-pointer", "-O1"]
if COV:
    CXXFLAGS += ["--coverage", "-O1"]
if TSAN:
    CXXFLAGS += ["-fsanitize=thread", "--param", "tsan-distinguish-volatile=1", "-fno-omit-frame-pointer", "-O1", "-DCUSIM_TSAN=1"]


def _match(text, i, open_ch, close_ch):
    """index just past the bracket that closes text[i] == open_ch"""
    depth = 0
    while i < len(text):
        c = text[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced")


def rewrite_launches(text, name):
    out, pos, n = [], 0, 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            break
        j = text.index(">>>", i)
        cfg = text[i + 3:j]
        # kernel expression in front of <<<: identifier with optional template arguments
        k = i
        while text[k - 1].isspace():
            k -= 1
        if text[k - 1] == ">":
            depth, k = 0, k - 1
            while True:
                if text[k] == ">":
                    depth += 1
                elif text[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                k -= 1
        while text[k - 1].isalnum() or text[k - 1] == "_":
            k -= 1
        kernel = text[k:i].strip()
        a0 = j + 3
        while text[a0].isspace():
            a0 += 1
        assert text[a0] == "(", (name, text[i - 40:i + 40])
        a1 = _match(text, a0, "(", ")")
        args = text[a0:a1]
        e = a1
        while text[e].isspace():
            e += 1
        assert text[e] == ";", (name, text[a1 - 40:a1 + 10])
        out.append(text[pos:k])
        out.append('cusim::launch("%s", cusim::cfg(%s), [&] { %s%s; })' % (kernel, cfg, kernel, args))
        pos = a1
        n += 1
    out.append(text[pos:])
    return "".join(out), n


def rewrite_asm(text, name):
    out, pos, n = [], 0, 0
    for m in re.finditer(r"\basm\s+volatile\s*\(", text):
        if m.start() < pos:
            continue
        end = _match(text, m.end() - 1, "(", ")")
        body = text[m.end():end - 1]
        out.append(text[pos:m.start()])
        if "%%globaltimer" in body:
            var = re.search(r'"=l"\((\w+)\)', body).group(1)
            out.append("%s = cusim::globaltimer()" % var)
        else:
            first = re.search(r'"([^"\\]*)', body).group(1).strip()[:60]
            out.append('cusim::ptx_unavailable("inline PTX (%s)")' % first.replace('"', "'"))
        pos = end
        n += 1
    out.append(text[pos:])
    return "".join(out), n


COOP_RE = re.compile(r"cudaLaunchCooperativeKernel\(deep \? \(void\*\)k_gn_persistent<4> : \(void\*\)k_gn_persistent<2>, dim3\(blocks\),\s*"
                     r"dim3\(kIcpThreads\), args, 0, L\.stream\)")
COOP_NEW = ("(deep ? cusim::launch_coop(\"k_gn_persistent<4>\", cusim::cfg(dim3(blocks), dim3(kIcpThreads)), "
            "[&] { k_gn_persistent<4>(kpv, jv, slots, ticket, pub, cd); }) : "
            "cusim::launch_coop(\"k_gn_persistent<2>\", cusim::cfg(dim3(blocks), dim3(kIcpThreads)), "
            "[&] { k_gn_persistent<2>(kpv, jv, slots, ticket, pub, cd); }))")


def generate():
    os.makedirs(GEN, exist_ok=True)
    counts = {}
    for src in SOURCES:
        text = open(os.path.join(CSRC, src)).read()
        text, nl = rewrite_launches(text, src)
        text, na = rewrite_asm(text, src)
        nc = 0
        if src == "sb_icp.cu":
            text, nc = COOP_RE.subn(COOP_NEW, text)
            assert nc == 1, "the cooperative launch site of sb_icp.cu changed"
            assert "void* args[] = {&kpv, &jv, &slots, &ticket, &pub, &cd};" in text, "argument list of the cooperative launch changed"
        assert "<<<" not in text and "asm volatile" not in text
        counts[src] = (nl, na, nc)
        dst = os.path.join(GEN, src.replace(".cu", ".cpp"))
        new = '#line 1 "%s"\n%s' % (os.path.join(CSRC, src), text)
        if not os.path.exists(dst) or open(dst).read() != new:
            open(dst, "w").write(new)
    return counts


def is_stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("cusim.hpp", "cusim_rt.cpp", "build_sim.py")]
    deps += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return OUT
    counts = generate()
    if verbose:
        print("rewrites (launches, asm, cooperative):", counts)
    shim = os.path.join(HERE, "shim")  # <cuda.h>, <cuda_runtime.h> -> cusim.hpp
    inc = ["-I", shim, "-I", HERE, "-I", CSRC, "-include", os.path.join(HERE, "cusim.hpp")]
    objs, procs = [], []
    for src in SOURCES + ["cusim_rt.cpp"]:
        cpp = os.path.join(HERE, src) if src == "cusim_rt.cpp" else os.path.join(GEN, src.replace(".cu", ".cpp"))
        obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
        flags = list(CXXFLAGS)
        if TSAN and src == "cusim_rt.cpp":  # the executor's own bookkeeping is not what is being checked
            flags = [f for f in flags if f != "-fsanitize=thread"]
        cmd = ["g++"] + flags + inc + ["-c", "-o", obj, cpp]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("g++ failed on %s:\n%s" % (src, out.decode()[-6000:]))
        if verbose and out.strip():
            print(out.decode()[-3000:])
    subprocess.check_call(["g++", "-shared", "-pthread"] + (["-fsanitize=address"] if ASAN else []) + (["-fsanitize=thread", "-Wl,-Bsymbolic-functions"] if TSAN else []) + (["--coverage"] if COV else []) + ["-o", OUT] + objs)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
