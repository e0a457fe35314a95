"""Summarises an ncu raw/source CSV export pair: key metrics, stall breakdown, hottest instructions."""
import csv, collections, sys
raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum [", "dram__bytes_write.sum [", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread [", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum [",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max [",
        "launch__grid_size", "launch__block_size", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum [", "l1tex__data_pipe_lsu_wavefronts.sum ["]
for i, h in enumerate(hdr):
    key = h + " ["
    if any(key.startswith(w) or h == w.strip(" [") for w in want):
        print("%-90s %-14s %s" % (h, units[i], [r[i] for r in data]))
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for s in stalls:
        try: tot[s] += float(r[col[s]])
        except: pass
allv = sum(tot.values()) or 1
print({k: round(100 * v / allv, 1) for k, v in tot.most_common(9)})
rs = []
for r in data:
    try: rs.append((int(r[col["# Samples"]]), r[col["Source"]][:70], {s[6:]: r[col[s]] for s in stalls if r[col[s]] not in ("0", "")}))
    except: pass
print("instructions", len(rs), "samples", sum(x[0] for x in rs))
ops = collections.Counter()
for r in data:
    try: ops[r[col["Source"]].split()[0]] += int(r[col["Instructions Executed"]])
    except: pass
print("executed by opcode:", ops.most_common(14))
rs.sort(key=lambda x: -x[0])
for x in rs[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]: print(x[0], x[1], x[2])
