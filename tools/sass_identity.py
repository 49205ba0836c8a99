"""Are the kernels of two builds the same machine code?  (development tool; needs cuobjdump and cu++filt, no GPU)

    git archive <commit> periodic_lbm_b200/csrc include | tar -x -C /tmp/old && make -C /tmp/old/periodic_lbm_b200/csrc -j8 <objects>
    python tools/sass_identity.py /tmp/old/periodic_lbm_b200/csrc periodic_lbm_b200/csrc plbm_lbm.o plbm_lbm2.o ...

Per object file: kernels (demangled names; the hash of the anonymous namespace differs from build to build) whose SASS -- every
instruction with its operands, encodings stripped -- is identical / differs / exists on one side only.  For the ones that differ, whether
at least the multiset of opcodes is the same (instruction order / register allocation only).  Used at the end of round 2 to show that
the kernels which the last GPU calls did not re-run are byte-for-byte the ones the whole-suite run had validated."""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs, cur, name = {}, [], None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = cur
            name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = []
        elif name is not None and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            cur.append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
    if name:
        funcs[name] = cur
    return funcs


def opcodes(lines):
    return collections.Counter(re.match(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln).group(1) for ln in lines)


if __name__ == "__main__":
    old_dir, new_dir, objs = sys.argv[1], sys.argv[2], sys.argv[3:]
    for obj in objs:
        a, b = kernels(f"{old_dir}/{obj}"), kernels(f"{new_dir}/{obj}")
        same = [k for k in a if k in b and a[k] == b[k]]
        diff = [k for k in a if k in b and a[k] != b[k]]
        reorder = [k for k in diff if opcodes(a[k]) == opcodes(b[k])]
        print(f"{obj}: {len(a)} old / {len(b)} new kernels; identical SASS {len(same)}; different {len(diff)} "
              f"(same opcode multiset: {len(reorder)}); only old {len([k for k in a if k not in b])}; only new {len([k for k in b if k not in a])}")
        for k in diff:
            print("    differs:", k[:160], "" if k in reorder else "  <-- opcode counts differ")
