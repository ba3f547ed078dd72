"""Static look at a kernel's SASS: every loop (backward branch) with its length and opcode mix.

    cuobjdump -sass k_rsa64.o | python tools/sass_loops.py 'rsa_verify_kernelILi64ELi4ELb0ELb1'
Used to keep the squaring variant's hot code inside the instruction cache and to see where ptxas puts register moves
(IMAD.MOV shares the FMA pipe with IMAD.WIDE)."""
import collections
import re
import sys


def main():
    want = sys.argv[1]
    name, ins = None, []
    for line in sys.stdin:
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name is None or want not in name:
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_to_i = {a: i for i, (a, _) in enumerate(ins)}
    print("instructions", len(ins), "bytes", len(ins) * 16)
    tot = collections.Counter()
    for i, (a, s) in enumerate(ins):
        op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0]
        tot[".".join(op.split(".")[:2])] += 1
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(?:`\([^)]*\)|0x([0-9a-f]+))", s)
        if m and m.group(1):
            t = int(m.group(1), 16)
            if t in addr_to_i and t <= a:
                j = addr_to_i[t]
                mix = collections.Counter()
                for _, b in ins[j:i + 1]:
                    o = re.sub(r"^@!?U?P\d+\s+", "", b).split()[0]
                    mix[".".join(o.split(".")[:2])] += 1
                print(f"loop {j:5d}..{i:5d} len {i - j + 1:5d}  " + "  ".join(f"{k}:{v}" for k, v in mix.most_common(7)))
    print("whole kernel: " + "  ".join(f"{k}:{v}" for k, v in tot.most_common(10)))


if __name__ == "__main__":
    main()
