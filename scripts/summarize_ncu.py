"""Turn the .ncu-rep files brought back in gpurun_out/ into the text summaries kept under
profiles/ (run here, no GPU needed):  python scripts/summarize_ncu.py <tag>"""
import collections, csv, json, pathlib, subprocess, sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["prof_tc2", "prof_stream2", "prof_compact2"]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'launch__grid_size', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max.per_second', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__waves_per_multiprocessor']


def to_bytes(val, unit):
    try:
        v = float(val)
    except ValueError:
        return None
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def raw_page(rep):
    res = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(res.stdout.splitlines()))
    return rows[0], rows[1], rows[2:]


traffic = {}
for name in names:
    rep = OUT / f"{name}.ncu-rep"
    if not rep.exists():
        continue
    hdr, units, rows = raw_page(rep)
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    out = PROF / f"{tag}_ncu_{name}.csv"
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows))])
        for i in idx:
            w.writerow([hdr[i], units[i]] + [r[i] for r in rows])
    ir, iw, it = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
    per = []
    for r in rows:
        rd, wr = to_bytes(r[ir], units[ir]), to_bytes(r[iw], units[iw])
        per.append({"kernel": r[hdr.index('Kernel Name')].split('(')[0], "dram_read_bytes": rd, "dram_write_bytes": wr,
                    "duration": r[it] + " " + units[it]})
    traffic[name] = per
    print("wrote", out)

launches = OUT / "launches_full.csv"
if launches.exists():
    lines = [l for l in launches.read_text().splitlines() if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        nm = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += v; total += v
    (PROF / f"{tag}_launches_full.csv").write_text(launches.read_text())
    with open(PROF / f"{tag}_launch_shares.md", "w") as fh:
        fh.write("| kernel | launches | total ms (ncu, serialised, cold) | share |\n|---|---|---|---|\n")
        for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{nm}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / total:.1f} % |\n")
    print((PROF / f"{tag}_launch_shares.md").read_text())
(PROF / f"{tag}_ncu_traffic.json").write_text(json.dumps(traffic, indent=1))
print(json.dumps(traffic, indent=1)[:3000])
