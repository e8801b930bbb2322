#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box through gpurun; outputs under gpurun_out/prof_$TAG/).  Every pass is bounded
# (a handful of kernel instances inside the NVTX range of ONE warmed-up step, own timeout): `ncu --set full` replays each
# kernel ~40 times, an unbounded pass over a whole step costs tens of GPU-minutes.
#   launches_bench_steps2.csv    every kernel launch of the bench command (2 timed steps), gpu__time_duration only
#   <group>.ncu-rep              ncu --set full of a few instances of one kernel group
TAG=${1:-r02}
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_steps2.json 2> $OUT/bench_steps2.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $OUT/launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
full() {   # name, workload, kernel regex, count
  timeout 200 ncu --set full --clock-control none --nvtx --nvtx-include "profiled/" -k regex:"$3" -c $4 -o $OUT/$1 \
      python scripts/profile_step.py $2 > $OUT/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
full knn pix3d "k_nn_grid|k_grid_build" 3
full gemm pix3d "k_gemm_tc|k_gemm_tn" 5
full gather pix3d "k_gather_fwd|k_gather_bwd" 4
full losses pix3d "k_normals|k_sample|k_normalize|k_cdf|k_areas|k_edge|k_chamfer_bwd|k_normal_loss" 10
full heads pix3d "k_head|k_texrows|k_skinny|k_pack_b2" 6
full cubify4 cubify4 "k_emit|k_faceflags|k_lattice|k_scan" 7
full align_shapenet shapenet_residual "k_proj_gather|k_map_to_rows|k_rows_to_map" 6
ls -la $OUT
