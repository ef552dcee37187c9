"""Development aid: condense an .ncu-rep into the few numbers DESIGN.md / profiles/ quote.
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...] > profiles/xyz.md
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_inst_executed_op_utcmma.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "smsp__inst_executed_op_ldgsts.sum",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max.per_second",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            print(f"## {path.split('/')[-1]} — {d.get('Kernel Name', '?')[:90]}")
            print()
            print("| metric | value | unit |")
            print("|---|---:|---|")
            for k in KEYS:
                cands = [name for name in hdr if (name == k or name.endswith("." + k)) and d.get(name, "") != ""]
                if cands:
                    name = min(cands, key=len)
                    print(f"| `{k}` | {d[name]} | {u[name]} |")
            print()


if __name__ == "__main__":
    main()
