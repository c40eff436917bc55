run() { echo -n "dbg=$1 mt=$2 res=$3 :: "; PG_PLANES_DEBUG=$1 PG_PLANES_MT=$2 PG_PLANES_RESIDENT=$3 timeout 120 python tools/conv_bench.py $4 $5 $6 $7 1 5 2; }
# skeleton 7; +no res ring 16391; +no tmem 49159; + no compute 114695 ; only no compute 65543; only no tmem 32775
for dbg in 7 16391 49159 114695 65543 32775; do run $dbg 4 1 32 3 1 1648800; done
for dbg in 7 16391 49159 114695; do run $dbg 2 0 128 3 1 412200; done
