#!/usr/bin/env python
"""Static check of a kernel's hot loop: compile ndt2d_kernels.cu (or another .cu) to a cubin, disassemble one function and
print, for every backward branch, the length of the loop it closes and the instruction mix inside it.
  python tools/looplen.py [-D...] [--fun SUBSTR] [--src FILE] [--dump]"""
import collections, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def main():
    args = sys.argv[1:]
    fun, src, dump, extra = "k_alignILi0ELb1ELb0ELb0", "ndt2d_kernels.cu", False, []
    while args:
        a = args.pop(0)
        if a == "--fun": fun = args.pop(0)
        elif a == "--src": src = args.pop(0)
        elif a == "--dump": dump = True
        else: extra.append(a)
    cubin = os.path.join(tempfile.gettempdir(), "looplen_%d.cubin" % os.getpid())
    cmd = ["nvcc", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
           "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "gtsam_ndt_b200", "csrc"), "-cubin", "-o", cubin,
           os.path.join(ROOT, "gtsam_ndt_b200", "csrc", src)] + extra
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode: sys.exit(r.stderr[-4000:])
    log = r.stderr.splitlines()
    for i, l in enumerate(log):
        if fun in l and "Compiling" in l:
            print("\n".join(x.strip() for x in log[i:i + 4] if "registers" in x or "spill" in x))
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    os.unlink(cubin)
    blocks = sass.split("Function : ")
    for b in blocks[1:]:
        name = b.split("\n", 1)[0]
        if fun not in name: continue
        ins = []
        for l in b.splitlines():
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
        print(name, len(ins), "instructions")
        addr_index = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr_index:
                j = addr_index[int(m.group(1), 16)]
                body = ins[j:i + 1]
                mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in body)
                print("loop 0x%04x..0x%04x: %d instructions  %s" % (ins[j][0], a, len(body), dict(mix.most_common())))
                if dump and len(body) > 60:
                    for _, x in body: print("    ", x)
if __name__ == "__main__":
    main()
