# usage: bash scripts/knn_sweep.sh  -- rebuilds chamfer.cu with a few parameter sets and times each (run under gpurun)
set -e
F=mesh_r-cnn_computer_vision_project_b200/csrc/chamfer.cu
cp $F /tmp/chamfer_orig.cu
for cfg in "256 64 1 48 32" "256 64 1 64 32" "256 64 1 96 32" "128 64 1 64 32" "256 32 1 64 32"; do
  set -- $cfg
  sed -e "s/constexpr int TILE = [0-9]*;/constexpr int TILE = $1;/" -e "s/constexpr int THREADS = [0-9]*;/constexpr int THREADS = $2;/" \
      -e "s/constexpr int QPT = [0-9]*;/constexpr int QPT = $3;/" -e "s/constexpr int QCAP = [0-9]*;/constexpr int QCAP = $4;/" \
      -e "s/constexpr int STEP = [0-9]*;/constexpr int STEP = $5;/" /tmp/chamfer_orig.cu > $F
  python -m meshrcnn_b200.build > /dev/null 2>&1 || { echo "build failed for $cfg"; continue; }
  echo "TILE=$1 THREADS=$2 QPT=$3 QCAP=$4 STEP=$5: $(timeout 100 python scripts/time_knn.py 10 2>&1 | tail -1)"
done
cp /tmp/chamfer_orig.cu $F
python -m meshrcnn_b200.build > /dev/null 2>&1
