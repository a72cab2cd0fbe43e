"""Static schedule of a kernel's SASS (developer tool; CPU only).
usage: sass_sched.py <object> <function-substring> [--dump]
Decodes the control word of every instruction (stall count, yield, write/read barrier, wait mask) from
`cuobjdump -sass` and prints, per basic block, the instruction count, the sum of the stall counts (the cycles a
lone warp needs to issue the block when no scoreboard wait bites) and the opcode mix."""
import re, subprocess, sys, collections

def load(obj, fun):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if fun in name:
            return name, b
    raise SystemExit("function not found")

def parse(body):
    ins = []
    lines = body.split("\n")
    i = 0
    pat = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/")
    pat2 = re.compile(r"^\s*/\* 0x([0-9a-f]{16}) \*/")
    while i < len(lines):
        m = pat.match(lines[i])
        if m and i + 1 < len(lines):
            m2 = pat2.match(lines[i + 1])
            if m2:
                addr = int(m.group(1), 16)
                text = m.group(2).strip()
                hi = int(m2.group(1), 16)
                ctrl = hi >> 41
                ins.append(dict(addr=addr, text=text, stall=ctrl & 0xf, yld=(ctrl >> 4) & 1, wbar=(ctrl >> 5) & 7,
                                rbar=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 0x3f))
                i += 2
                continue
        i += 1
    return ins

def opcode(text):
    t = text
    if t.startswith("@"):
        t = t.split(" ", 1)[1]
    return t.split(" ")[0].split(".")[0]

if __name__ == "__main__":
    name, body = load(sys.argv[1], sys.argv[2])
    ins = parse(body)
    print(name, len(ins), "instructions")
    targets = set()
    for k in ins:
        m = re.search(r"\b(BRA|BSSY|CALL)\S*\s.*?`?\(?(0x[0-9a-f]+)\)?", k["text"])
        if k["text"].find("BRA") >= 0 or k["text"].find("BSSY") >= 0:
            mm = re.search(r"0x([0-9a-f]+)\s*$", k["text"].rstrip(" ;"))
            if mm:
                targets.add(int(mm.group(1), 16))
    # basic blocks
    start = 0
    blocks = []
    for idx, k in enumerate(ins):
        op = opcode(k["text"])
        if k["addr"] in targets and idx > start:
            blocks.append((start, idx)); start = idx
        if op in ("BRA", "EXIT", "RET", "BRX", "JMP") or op == "BSYNC" or op == "WARPSYNC":
            blocks.append((start, idx + 1)); start = idx + 1
    if start < len(ins):
        blocks.append((start, len(ins)))
    dump = "--dump" in sys.argv
    for (a, b) in blocks:
        seg = ins[a:b]
        st = sum(max(1, k["stall"]) for k in seg)
        mix = collections.Counter(opcode(k["text"]) for k in seg)
        waits = sum(1 for k in seg if k["wait"])
        print("block %05x..%05x  n=%4d  stall_sum=%5d  cyc/inst=%.2f  waits=%3d  %s  | last: %s" % (
            seg[0]["addr"], seg[-1]["addr"], len(seg), st, st / len(seg), waits,
            " ".join("%s:%d" % kv for kv in mix.most_common(6)), seg[-1]["text"][:60]))
        if dump:
            for k in seg:
                print("   %05x s%2d %s w%d r%d wait%02x  %s" % (k["addr"], k["stall"], "Y" if k["yld"] else " ", k["wbar"], k["rbar"], k["wait"], k["text"]))
