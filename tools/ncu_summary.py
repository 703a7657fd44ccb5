#!/usr/bin/env python
"""One-page summary of an .ncu-rep (runs `ncu -i ... --page raw --csv` here, no GPU needed): per captured kernel the duration, DRAM /
L2 / L1 traffic and utilisation, instruction counts and issue utilisation, occupancy.  usage: tools/ncu_summary.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'registers/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem/block'), ('launch__shared_mem_per_block_static', 'static smem/block'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('dram__sectors_read.sum', 'DRAM sectors read'),
    ('lts__t_bytes.sum', 'L2 bytes'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 sector hit rate %'),
    ('l1tex__t_bytes.sum', 'L1 bytes'), ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/TEX throughput % of peak'),
    ('l1tex__data_pipe_lsu_wavefronts.sum', 'L1 data-pipe wavefronts'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'L1 data-pipe wavefronts % of peak'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', '... of which shared memory'),
    ('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'global load requests'),
    ('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'global load sectors'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 sector hit rate %'),
    ('smsp__inst_executed.sum', 'warp instructions'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__inst_executed.avg.per_cycle_elapsed', 'IPC per SM'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput % of peak'),
    ('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU pipe %'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe %'),
    ('sm__cycles_elapsed.avg', 'SM cycles'),
    ('smsp__average_warp_latency_per_inst_issued.ratio', 'warp cycles per issued instruction'),
    ('smsp__warps_eligible.avg.per_cycle_active', 'eligible warps per scheduler'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"### `{r[col['Kernel Name']]}`\n\n```")
        for key, label in WANT:
            if key in col and r[col[key]] not in ('', 'n/a'):
                print(f'{label:42s} {r[col[key]]:>18s} {units[col[key]]}')
        print('```\n')


if __name__ == '__main__':
    main(sys.argv[1])
