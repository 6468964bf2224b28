"""tools/sass_compare.py BEFORE_DIR AFTER_DIR -- are the plain fused kernels byte-for-byte the same code?

Each directory holds `cuobjdump -sass` dumps (*.sass, addresses and encodings stripped) of the fused_f*.o objects of one
build.  Functions are matched in order after dropping the extended-I/O instantiations (template argument EX = true);
used to prove that adding the EX code path left every default kernel's instruction stream untouched.
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
    print(f"{same} of {total} plain kernels have identical SASS")
    if ex_sizes:
        ex_sizes.sort()
        print(f"{len(ex_sizes)} extended-I/O kernels: {ex_sizes[0]} / {ex_sizes[len(ex_sizes) // 2]} / {ex_sizes[-1]} instructions (min / median / max)")
    return 0 if same == total else 1


if __name__ == "__main__":
    sys.exit(main())
