"""Dev tool: turn the raw ncu CSVs in gpurun_out/ into the committed summaries under profiles/.
usage: python tools/summarize_profiles.py <suffix e.g. _b> <round tag e.g. r01b>"""
import collections, csv, io, json, subprocess, sys

suf, tag = sys.argv[1], sys.argv[2]
TIME = {'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}
BYTE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}

def rows(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    return list(csv.DictReader(io.StringIO(''.join(lines))))

# 1. launch list
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows(f'gpurun_out/launches_v2_99{suf}.csv'):
    n = r['Kernel Name'].split('(')[0].replace('dd3d::<unnamed>::', '').replace('void ', '')[:48]
    tot[n] += float(r['Metric Value'].replace(',', '')) * TIME.get(r['Metric Unit'], 1e-6)
    cnt[n] += 1
s = sum(tot.values())
out = [f"# {tag}: ncu launch list, V2-99 DD3D bf16 B=32 900x1600 (5 forwards: warm-up+step device path, warm-up+step host path, 1 profiled)",
       "# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv python bench.py --workload v2_99 --steps 1 --warmup 1 --cpu-images 0",
       "# per-launch times are cold-cache and serialised: compare SHARES with bench.py's live kernels_ms_per_step, not absolutes",
       "kernel,launches,total_ms,share_pct"]
for n, v in sorted(tot.items(), key=lambda x: -x[1]):
    out.append(f"{n},{cnt[n]},{v:.3f},{100 * v / s:.2f}")
open(f'profiles/{tag}_launches_v2_99_summary.csv', 'w').write('\n'.join(out) + '\n')

# 2. per conv launch
by = collections.defaultdict(dict)
for r in rows(f'gpurun_out/conv_metrics_v2_99{suf}.csv'):
    try:
        v = float(r['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    m, u = r['Metric Name'], r['Metric Unit']
    if m.startswith('dram__bytes'):
        v *= BYTE.get(u, 1)
    if m.startswith('gpu__time'):
        v *= TIME.get(u, 1)
    by[int(r['ID'])][m] = v
names = (["stem_2 64->64 k3", "stem_3 64->128 k3s2"] + [f"OSA2_1 {x}" for x in ["l0", "l1", "l2", "l3", "l4", "concat 768->256"]] +
         [f"OSA3_{b} {x}" for b in (1, 2, 3) for x in ["l0", "l1", "l2", "l3", "l4", "concat ->512"]] +
         [f"OSA4_{b} {x}" for b in range(1, 10) for x in ["l0", "l1", "l2", "l3", "l4", "concat ->768"]] +
         [f"OSA5_{b} {x}" for b in (1, 2, 3) for x in ["l0", "l1", "l2", "l3", "l4", "concat ->1024"]] +
         ["fpn_lateral5", "fpn_output5", "fpn_lateral4", "fpn_output4", "fpn_lateral3", "fpn_output3", "fpn_lateral2", "fpn_output2", "top_block.p6"] +
         [f"{t}_tower.{i} (5 levels)" for t in ("cls", "box2d", "box3d") for i in range(4)] +
         ["cls_logits N=16 f32 (5 levels)", "box2d_reg+centerness N=16 f32 (5 levels)", "box3d_all N=112 f32 (5 levels)"])
o = [f"# {tag}: every conv_igemm launch of ONE V2-99 B=32 forward (2nd forward): duration, tensor-pipe %, DRAM traffic",
     "# command: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_igemm -s 122 -c 122",
     "idx,layer,ms,tensor_pipe_pct,sm_throughput_pct,dram_read_MB,dram_write_MB,dram_GBs"]
tt = tb = tw = 0
TP = 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'
for k, i in enumerate(sorted(by)):
    d = by[i]; t = d['gpu__time_duration.sum']; rd = d['dram__bytes_read.sum']; wr = d['dram__bytes_write.sum']
    tt += t; tb += rd + wr; tw += t * d[TP]
    o.append(f"{k},{names[k]},{t:.4f},{d[TP]:.1f},{d['sm__throughput.avg.pct_of_peak_sustained_elapsed']:.1f},{rd / 1e6:.1f},{wr / 1e6:.1f},{(rd + wr) / 1e9 / (t / 1e3):.0f}")
o.append(f"# total: {tt:.2f} ms, DRAM traffic {tb / 1e9:.2f} GB per forward, time-weighted tensor pipe {tw / tt:.1f} %")
open(f'profiles/{tag}_conv_launches_v2_99.csv', 'w').write('\n'.join(o) + '\n')
json.dump({"v2_99": tb, "note": f"sum of dram__bytes_read.sum + dram__bytes_write.sum over the 122 conv_igemm launches of one B=32 forward (profiles/{tag}_conv_launches_v2_99.csv)"},
          open('profiles/conv_igemm_traffic.json', 'w'), indent=1)
print(o[-1])

# 3. full capture of the tower conv
raw = subprocess.run(['ncu', '-i', f'gpurun_out/prof_tower{suf}.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
keys = ['gpu__time_duration.sum', TP, 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.avg.per_second']
md = [f"# {tag} -- ncu --set full capture of the FCOS tower conv (conv_igemm_kernel<halo>, 3x3 256->256, 5 FPN levels per launch, B=8)", "",
      "`ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 229 -c 2 -o gpurun_out/prof_tower python bench.py --workload v2_99 --batch 8 --steps 1 --warmup 1 --cpu-images 0`", "",
      "| metric | " + " | ".join(f"launch {i + 1}" for i in range(len(rr) - 2)) + " |", "|---|" + "---|" * (len(rr) - 2)]
for k in keys:
    if k in hdr:
        j = hdr.index(k)
        md.append(f"| {k} [{units[j]}] | " + " | ".join(r[j] for r in rr[2:]) + " |")
open(f'profiles/{tag}_prof_tower_summary.md', 'w').write('\n'.join(md) + '\n')
print('\n'.join(md[-14:]))
