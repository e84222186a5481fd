"""Decode the scheduling control words of a cuobjdump -sass listing (sm_100a, 128-bit instructions).

    cuobjdump -sass -fun <mangled> file.o > k.sass
    python tools/sass_sched.py k.sass [first_addr last_addr]

Prints, per instruction: address, stall count, yield, write/read scoreboard, wait mask, text.
With a range, also the sum of the static stall counts (the cycles one warp ALONE needs to issue the
range when no scoreboard wait blocks) and the instruction mix.  Control field (bits 105..125 of the
instruction): stall[4] yield[1] wbar[3] rbar[3] wait[6] reuse[4].
"""
import re
import sys
from collections import Counter

INS = re.compile(r"^\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/")
HI = re.compile(r"^\s*/\* 0x([0-9a-f]{16}) \*/")


def parse(path):
    out = []
    cur = None
    for line in open(path):
        m = INS.match(line)
        if m:
            cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
            continue
        m = HI.match(line)
        if m and cur is not None:
            cur[3] = int(m.group(1), 16)
            out.append(cur)
            cur = None
    return out


def ctrl(hi):
    c = hi >> 41
    return dict(stall=c & 15, yld=(c >> 4) & 1, wbar=(c >> 5) & 7, rbar=(c >> 8) & 7,
                wait=(c >> 11) & 63, reuse=(c >> 17) & 15)


def main():
    ins = parse(sys.argv[1])
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
    total = 0
    mix = Counter()
    for addr, text, _, hw in ins:
        if not (lo <= addr <= hi):
            continue
        c = ctrl(hw)
        total += c["stall"]
        op = text.split()[0] if not text.startswith("@") else text.split()[1]
        mix[op.split(".")[0]] += 1
        wb = c["wbar"] if c["wbar"] != 7 else "-"
        rb = c["rbar"] if c["rbar"] != 7 else "-"
        print(f"{addr:05x} st={c['stall']:2d} y={c['yld']} w={wb} r={rb} wait={c['wait']:06b}  {text}")
    print(f"# instructions {sum(mix.values())}  static stall sum {total}")
    print("# mix", dict(mix.most_common()))


if __name__ == "__main__":
    main()
