#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box through gpurun; outputs under gpurun_out/prof_$TAG/).  Every pass is bounded
# (a handful of kernel instances of ONE warmed-up step between cudaProfilerStart/Stop, own timeout): `ncu --set full`
# replays each kernel ~40 times, an unbounded pass over a whole step costs tens of GPU-minutes.
#   usage: capture_profiles.sh TAG [group ...]      (no group = all)
#   launches_one_step.csv        every kernel launch of one step (forward + backward), gpu__time_duration only
#   <group>.ncu-rep              ncu --set full of a few instances of one kernel group
TAG=${1:-r02}; shift
GROUPS_WANTED="$*"
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
want() { [ -z "$GROUPS_WANTED" ] || [[ " $GROUPS_WANTED " == *" $1 "* ]]; }
if want launches; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file $OUT/launches_one_step.csv python scripts/profile_step.py pix3d > $OUT/launches.log 2>&1
  echo "launches rc=$?"
fi
full() {   # name, workload, kernel regex, count
  want $1 || return 0
  timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:"$3" -c $4 -o $OUT/$1 \
      python scripts/profile_step.py $2 > $OUT/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
full knn pix3d "k_nn_grid|k_grid_build" 3
full gemm pix3d "k_gemm_tc" 4
full wgrad pix3d "k_gemm_tn" 3
full gather pix3d "k_gather_fwd" 4
full gather_bwd pix3d "k_gather_bwd" 3
full losses pix3d "k_normals_fwd|k_sample$|k_normalize|k_cdf|k_areas|k_edge_fwd|k_normal_loss$|k_sum" 10
full losses_bwd pix3d "k_normals_bwd|k_sample_bwd|k_edge_bwd|k_chamfer_bwd|k_normal_loss_bwd" 7
full heads pix3d "k_head|k_texrows|k_skinny|k_pack_b2" 8
full cubify4 cubify4 "k_emit|k_faceflags|k_lattice|k_scan" 7
full align_shapenet shapenet_residual "k_proj_gather|k_map_to_rows|k_rows_to_map" 6
ls -la $OUT
