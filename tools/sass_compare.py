"""tools/sass_compare.py BEFORE_DIR AFTER_DIR -- are the plain fused kernels byte-for-byte the same code?

Each directory holds `cuobjdump -sass` dumps (*.sass, addresses and encodings stripped) of the objects of one build.
fused_f*.o: functions are matched in order after dropping the extended-I/O instantiations (template argument EX = true,
which also changed the mangled names); every other object: matched by name.  Used to prove that adding the EX code path
and the CPU-emulation hooks (SSFFT_EMUL, SSFFT_DYNAMIC_SMEM) left every default kernel's instruction stream untouched:
round 1 ended with 116 of 116 fused and 130 of 130 other kernels identical to the build before either existed.
    for f in fused_f32_a ...; do cuobjdump -sass build/$f.o | grep -E "^\\s+/\\*[0-9a-f]{4,}\\*/|Function :" \\
        | sed -E 's|/\\*[0-9a-f]{4,}\\*/||; s|/\\* 0x[0-9a-f]+ \\*/||' > DIR/$f.sass; done
"""
import glob
import os
import re
import sys


def load(path):
    funcs, cur = [], None
    for line in open(path):
        if "Function :" in line:
            cur = [line.split("Function :")[1].strip(), []]
            funcs.append(cur)
        elif cur:
            cur[1].append(line.strip())
    return funcs


def main():
    before, after = sys.argv[1], sys.argv[2]
    total = same = 0
    ex_sizes = []
    for f in sorted(glob.glob(os.path.join(before, "*.sass"))):
        a, b = load(f), load(os.path.join(after, os.path.basename(f)))
        if not any("fused_fft_kernel" in x[0] for x in a):
            # other objects: kernels keep their mangled names, match by name; kernels that only exist after are new
            bd = dict((k, i) for k, i in b)
            for ka, ia in a:
                total += 1
                if bd.get(ka) == ia:
                    same += 1
                else:
                    print(os.path.basename(f), "DIFFERENT" if ka in bd else "MISSING", ka[:120])
            continue
        ex_sizes += [len(x[1]) for x in b if "Lb0ELb1E" in x[0]]
        b = [x for x in b if "Lb0ELb1E" not in x[0]]
        assert len(a) == len(b), (f, len(a), len(b))
        for (ka, ia), (kb, ib) in zip(a, b):
            total += 1
            assert ka.split("EEELb")[0] == kb.split("EEELb")[0], (ka, kb)  # same configuration, same MOD flag follows
            if ia == ib:
                same += 1
            else:
                print(os.path.basename(f), "DIFFERENT", len(ia), len(ib), ka[:120])
    print(f"{same} of {total} kernels have identical SASS")
    if ex_sizes:
        ex_sizes.sort()
        print(f"{len(ex_sizes)} extended-I/O kernels: {ex_sizes[0]} / {ex_sizes[len(ex_sizes) // 2]} / {ex_sizes[-1]} instructions (min / median / max)")
    return 0 if same == total else 1


if __name__ == "__main__":
    sys.exit(main())
