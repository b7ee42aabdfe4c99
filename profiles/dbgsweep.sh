for shape in "16 16 3 224" "32 16 3 224" "64 64 3 56" "256 256 3 14"; do
for op in 0 1; do
for d in 0 1 2 4 6 8 16 22 3 11 9 27 31; do
  set -- $shape
  echo -n "shape=($shape) op=$op dbg=$d : "
  HPFG_TC_DBG=$d python profiles/layer_bench.py 32 $op $1 $2 $3 $4 30
done; done; done
