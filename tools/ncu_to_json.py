"""Condense .ncu-rep files into profiles/ncu_latest.json (what bench.py quotes as roofline.traffic / roofline.ncu).
usage: python tools/ncu_to_json.py cfg2=gpurun_out/x_fused_cfg2.ncu-rep cfg4=... > profiles/ncu_latest.json"""
import csv, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
KEYS = {'gpu__time_duration.sum': 'duration', 'smsp__inst_executed.sum': 'warp_instructions',
        'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
        'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
        'l1tex__t_sector_hit_rate.pct': 'l1_hit_pct', 'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
        'launch__registers_per_thread': 'registers', 'launch__grid_size': 'grid',
        'sm__inst_executed_pipe_fma.sum': 'pipe_fma', 'sm__inst_executed_pipe_alu.sum': 'pipe_alu',
        'sm__inst_executed_pipe_xu.sum': 'pipe_xu', 'sm__inst_executed_pipe_lsu.sum': 'pipe_lsu'}
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 'msecond': 1e3, 'usecond': 1.0, 'nsecond': 1e-3}
out = {'csrc_sha': bench.csrc_sha()}     # bench.py quotes the traffic only while the kernel sources are the ones captured
for arg in sys.argv[1:]:
    name, path = arg.split('=')
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = {'kernel': r[hdr.index('Kernel Name')], 'report': path.split('/')[-1]}
    for k, short in KEYS.items():
        if k in hdr:
            v = float(r[hdr.index(k)].replace(',', ''))
            u = units[hdr.index(k)]
            if short in ('dram_read', 'dram_write'):
                v *= SCALE.get(u, 1.0); short_u = short + '_bytes'
            elif short == 'duration':
                v *= SCALE.get(u, 1.0); short_u = 'duration_us'
            else:
                short_u = short
            d[short_u] = v
    d['dram_bytes'] = d.get('dram_read_bytes', 0) + d.get('dram_write_bytes', 0)
    out[name] = d
print(json.dumps(out, indent=1))
