"""Dynamic opcode mix of one kernel from an ncu report (developer tool): ncu_opmix.py report.ncu-rep warps_x_iters"""
import collections, csv, io, subprocess, sys
rep, per = sys.argv[1], float(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; data = rows[2:]
iA, iE, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot, samp = collections.Counter(), collections.Counter()
for r in data:
    f = r[iA].split()
    op = f[1] if f[0].startswith('@') else f[0]
    op = op.split('.')[0].rstrip(';')
    tot[op] += int(r[iE]); samp[op] += int(r[iS])
T, S = sum(tot.values()), sum(samp.values())
print("warp instructions per warp-iteration: %.1f   (stall samples %d)" % (T / per, S))
for op, c in tot.most_common(28):
    print("%-10s %8.1f /iter  %5.1f%%  samples %5.1f%%" % (op, c / per, 100 * c / T, 100 * samp[op] / S))
