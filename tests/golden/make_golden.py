"""Generate the golden fixtures under tests/golden/ by running the compiled, unmodified reference
(oracle/_ref/ref_harness, built by `make -C oracle` from /root/reference/source).  Run in the build container:

    python tests/golden/make_golden.py

The reference has no golden vectors for this path; these dumps are its own outputs on
small instances of the five BASELINE.json configs.  Format: tests/refdump.py.  Files are xz-compressed.
"""
import lzma
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

CASES = {
    # cfg1: 2D advection k=2 (N=7 in BASELINE; fixture at N=4), RK3 on the assembled operator + sweep form
    "cfg1_adv_d2_k2_n4": "--dim 2 --nmax 4 --pa 2 --pl 3 --run grid,advection,rhs,stage --flux burgers --dump-tables 1 --dt 0.002",
    # cfg2: Lagrange round trip d=4 k=3 m=3 (N=8 in BASELINE; fixture at N=3)
    "cfg2_rt_d4_k3_n3": "--dim 4 --nmax 3 --pa 3 --pl 3 --run grid,roundtrip,der --dump-tables 1",
    # cfg3: 3D wave k=2 (N=7 in BASELINE; fixture at N=3)
    "cfg3_wave_d3_k2_n3": "--dim 3 --nmax 3 --pa 2 --pl 3 --run grid,wave,roundtrip --dump-tables 1 --dt 0.0005",
    # cfg4: 2D Burgers stage, Lagrange (m=3) and Hermite (m=3) flux interpolation, k=2 (fixture at N=4)
    "cfg4_burgers_lagr_d2_k2_n4": "--dim 2 --nmax 4 --pa 2 --pl 3 --run grid,rhs,stage --flux burgers1 --dump-tables 1 --dt 0.002",
    "cfg4_burgers_herm_d2_k2_n4": "--dim 2 --nmax 4 --pa 2 --ph 3 --intp herm --run grid,rhs,stage,roundtrip --flux burgers1 --dump-tables 1 --dt 0.002",
    "kpp_lagr_d2_k1_n4": "--dim 2 --nmax 4 --pa 1 --pl 2 --run grid,rhs --flux kpp --dump-tables 1",
    # cfg5: 6D Vlasov k=1 m=2 (N=7 in BASELINE; fixture at N=2), generalised 2D2V point-wise products
    "cfg5_vlasov_d6_k1_n2": "--dim 6 --nmax 2 --pa 1 --pl 2 --run grid,rhs,stage --flux vlasov --dump-tables 1 --dt 0.001",
    # 4D Vlasov 2D2V-like with the shipped degrees (k=3, m=4, mesh case 2), N=2
    "vlasov_d4_k3_m4_n2": "--dim 4 --nmax 2 --pa 3 --pl 4 --msh-lagr 2 --run grid,rhs --flux vlasov --dump-tables 1",
    # Vlasov coupling: velocity moments of f into the field solution (compute_moment_1D2V / _2D2V); no tables needed
    "moment_d3_k2_n4": "--dim 3 --nmax 4 --pa 2 --pl 3 --run grid,moment",
    "moment_d4_k1_n3": "--dim 4 --nmax 3 --pa 1 --pl 2 --run grid,moment",
    # full (non-sparse) grid and a 1D grid: edge cases of the schedule (d=1 is a single full sweep)
    "full_d2_k2_n3": "--dim 2 --nmax 3 --sparse 0 --pa 2 --pl 3 --run grid,rhs,roundtrip --flux linear --dump-tables 1",
    # adaptive (irregular) grids produced by DGAdapt::refine / coarsen: fibres are arbitrary subsets of the 1D tree
    "adapt_d2_k2_n6": "--dim 2 --nmax 6 --n0 2 --pa 2 --pl 3 --adapt-eps 0.02 --adapt-eta 0.012 --adapt-rounds 5 --run grid,rhs,roundtrip,stage --flux burgers --dump-tables 1 --dt 0.001",
    "adapt_d3_k1_n4": "--dim 3 --nmax 4 --n0 1 --pa 1 --pl 2 --adapt-eps 0.05 --adapt-eta 0.03 --adapt-rounds 4 --run grid,rhs,roundtrip --flux kpp --dump-tables 1",
    # the other FastRHS compositions: SameFlux / DiffFlux (DIM == 2) and SourceFastLagr, Lagrange and Hermite; 3-component source
    "variants_lagr_d2_k2_n4": "--dim 2 --nmax 4 --pa 2 --pl 3 --run grid,variants --flux kpp --dump-tables 1",
    "variants_herm_d2_k2_n4": "--dim 2 --nmax 4 --pa 2 --ph 3 --intp herm --run grid,variants --flux burgers --dump-tables 1",
    "variants_lagr_d3_k1_n3": "--dim 3 --nmax 3 --pa 1 --pl 2 --vecnum 2 --run grid,variants --flux kpp --dump-tables 1",
    # DiffusionRHS, FastRHSHamiltonJacobi, the *_coarse_grid transforms (mesh_nmax = N-1, N-2) and DGAdapt::indicator_norm
    "f4_lagr_d2_k2_n4": "--dim 2 --nmax 4 --pa 2 --pl 3 --run grid,f4 --flux kpp --dump-tables 2",
    "f4_lagr_d3_k1_n3": "--dim 3 --nmax 3 --pa 1 --pl 2 --run grid,f4 --flux burgers --dump-tables 2",
    # point-wise kernels beyond a scalar flux: a coupled two-variable flux through eval_fp_Lag, a coefficient of position through
    # var_coeff_u_Lagr_fast, the 2D2V Vlasov body with a second solution's field values broadcast by copy_up_intp_to_f
    "pw_d2_k2_n4_v2": "--dim 2 --nmax 4 --pa 2 --pl 3 --vecnum 2 --run grid,pw --dump-tables 1",
    "pw_vlasov_d4_k1_n3_v2": "--dim 4 --nmax 3 --pa 1 --pl 2 --vecnum 2 --run grid,pw --dump-tables 1",
    # one RK3SSP step of the coupled 2D2V Vlasov-Ampere system (f: interp_Vlasov_2D2V + rhs + penalty; E_t = -J by compute_moment_2D2V), per stage
    "vlasov_ampere_d4_k1_n3_v2": "--dim 4 --nmax 3 --pa 1 --pl 2 --vecnum 2 --run grid,vlasov_ampere --dump-tables 1 --dt 0.002",
    # 1D tables under the other boundary types of Basis::product_edge_dis_v / _u (table-generation parity only)
    "tables_bc_zero_k2_n4": "--dim 1 --nmax 4 --pa 2 --pl 3 --ph 3 --boundary zero --run grid --dump-tables 2",
    "tables_bc_inside_k2_n4": "--dim 1 --nmax 4 --pa 2 --pl 3 --ph 3 --boundary inside --run grid --dump-tables 2",
    # coefficient x gradient (var_coeff_gradu_Lagr_fast), source terms (source_from_lagr_to_rhs), the 1D2V Vlasov-Maxwell body of the shipped
    # example/07_vlasov_maxwell_sparse.cpp (interp_Vlasov_1D2V with B3, E1, E2 broadcast from a second solution); k = 2, m = 3 as shipped
    "pw2_vm_d3_k2_n3_v3": "--dim 3 --nmax 3 --pa 2 --pl 3 --vecnum 3 --run grid,pw2 --dump-tables 1",
    "line_d1_k2_n5": "--dim 1 --nmax 5 --pa 2 --pl 3 --run grid,rhs,roundtrip --flux burgers --dump-tables 1",
}


def main():
    only = sys.argv[1:]
    for name, args in CASES.items():
        if only and name not in only:
            continue
        tmp = os.path.join(HERE, name + ".dump")
        subprocess.run([HARNESS] + args.split() + ["--out", tmp, "--threads", "1"], check=True)
        with open(tmp, "rb") as f:
            raw = f.read()
        with lzma.open(tmp + ".xz", "wb", preset=9) as f:
            f.write(raw)
        os.remove(tmp)
        print(name, len(raw), "->", os.path.getsize(tmp + ".xz"))


if __name__ == "__main__":
    main()
