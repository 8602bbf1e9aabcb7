"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep, scratch) into small JSON files under profiles/ (tracked).

    python tools/ncu_summary.py gpurun_out/r02_prof_sweep.ncu-rep profiles/r02_sweep_ncu_summary.json --rows 1000000 --dim 512 --note "..."
"""
import argparse
import csv
import io
import json
import subprocess

KEYS = {
    'gpu__time_duration.sum': 'gpu_time',
    'dram__bytes_read.sum': 'dram_bytes_read',
    'dram__bytes_write.sum': 'dram_bytes_write',
    'lts__t_sector_hit_rate.pct': 'l2_hit_rate_pct',
    'launch__registers_per_thread': 'registers_per_thread',
    'sm__cycles_elapsed.avg.per_second': 'sm_clock',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct',
    'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed': 'memory_throughput_pct',
    'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed': 'tensor_pipe_active_pct',
    'sm__issue_active.avg.pct_of_peak_sustained_elapsed': 'issue_active_pct',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'xu_pipe_pct',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active': 'alu_pipe_pct',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'fma_pipe_pct',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active': 'lsu_pipe_pct',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed': 'shared_memory_wavefronts_pct',
    'smsp__inst_executed.sum': 'warp_instructions',
}
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0, 'Ghz': 1e9, 'Mhz': 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('out')
    ap.add_argument('--rows', type=int, default=0)
    ap.add_argument('--dim', type=int, default=0)
    ap.add_argument('--note', default='')
    ap.add_argument('--source', default='')
    args = ap.parse_args()
    txt = subprocess.run(['ncu', '-i', args.report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, d = rows[0], rows[1], rows[2]
    rec = dict(zip(hdr, d))
    out = dict(source=args.source, note=args.note, rows=args.rows, dim=args.dim, kernel=rec.get('Kernel Name'), grid=rec.get('Grid Size'),
               block=rec.get('Block Size'))
    for k, name in KEYS.items():
        if k in rec and rec[k] not in ('', 'n/a'):
            v = float(rec[k].replace(',', ''))
            u = units[hdr.index(k)]
            if u in SCALE:
                v *= SCALE[u]
                u = {'Gbyte': 'byte', 'Mbyte': 'byte', 'Kbyte': 'byte', 'ms': 's', 'us': 's', 'ns': 's', 'Ghz': 'hz', 'Mhz': 'hz'}.get(u, u)
            out[name] = v
            out[name + '_unit'] = u
    stalls = []
    for h in hdr:
        if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            try:
                stalls.append((float(rec[h].replace(',', '')), h.split('issue_stalled_')[-1]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    out['top_stall_samples'] = [dict(reason=n, samples=v) for v, n in stalls[:6]]
    if 'dram_bytes_read' in out:
        out['traffic_bytes_per_launch'] = out['dram_bytes_read'] + out.get('dram_bytes_write', 0.0)
        if out.get('gpu_time'):
            out['dram_gbs'] = out['traffic_bytes_per_launch'] / out['gpu_time'] / 1e9
    with open(args.out, 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out)[:600])


if __name__ == '__main__':
    main()
