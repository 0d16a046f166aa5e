"""Reduce an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) to the per-kernel table committed
under profiles/ and to profiles/r1/ncu_traffic.json (DRAM bytes per launch, read by bench.py's roofline object).

    ncu -i gpurun_out/<tag>/prof.ncu-rep --page raw --csv > raw.csv
    python profiles/extract_ncu.py raw.csv profiles/r1/ncu_full_c2.csv profiles/r1/ncu_traffic.json
"""
import csv, json, sys

WANT = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]
# bench.py phase name <- kernel name pattern (first match wins; GEMM launches are told apart by their order inside a step)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(raw, out_csv, out_json):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(out_csv, "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i][:100] for i in idx])
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    per = {}
    gemm_seen = 0
    ps_seen = 0
    for r in data:
        b = float(r[ir]) * UNIT.get(units[ir], 1.0) + float(r[iw]) * UNIT.get(units[iw], 1.0)
        name = r[ik]
        if "knm_umma_kernel" in name:
            key = "kmat_knm"
        elif "umma_gemm_ps_kernel" in name:
            # round 2: the two products with a pre-split right operand; inside one pipelined step V X^T (main stream) precedes V = K_nm L^-T
            key = ["gemm_v_sigma", "gemm_v"][ps_seen % 2]
            ps_seen += 1
        elif "umma_gram_tn_kernel" in name:
            key = "gemm_gram"          # Gram product straight from V (single-latent steps)
        elif "umma_gemm_nt_kernel" in name:
            if ps_seen:
                key = "gemm_gram"      # with the pre-split kernel in use, the first-generation kernel only runs the Gram product
            else:
                key = ["gemm_v_sigma", "gemm_gram", "gemm_v"][gemm_seen % 3]   # order of the three GEMMs inside one pipelined step
            gemm_seen += 1
        elif "tail2_step_kernel" in name or "tail_step_kernel" in name:
            key = "tail_step"
        elif "potf2_first" in name:
            key = "tail_potf2_first"
        elif "combine_kernel" in name:
            key = "combine_eta"
        else:
            continue
        per.setdefault(key, []).append(b)
    json.dump({k: sum(v) / len(v) for k, v in per.items()}, open(out_json, "w"), indent=1)
    print(json.dumps({k: round(sum(v) / len(v)) for k, v in per.items()}))


if __name__ == "__main__":
    main(*sys.argv[1:4])
