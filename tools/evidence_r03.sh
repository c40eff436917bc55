#!/bin/bash
# Round-3 evidence run (one B200, under gpurun): ncu launch list of one ragged decode of the bench clip, ncu --set full of
# the dominant conv, the tcgen05 attention and the fused upsampler, compute-sanitizer memcheck, run-to-run determinism,
# bench lines of the other BASELINE configs.  Outputs under gpurun_out/ (summaries are copied to profiles/ by hand).
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --launch-skip 184 --launch-count 184 --csv --log-file $O/r03_launches_decode.csv python tools/ncu_decode.py 2 > $O/r03_ncu_decode.log 2>&1
tail -1 $O/r03_ncu_decode.log
D="--set full --import-source on --kernel-name-base demangled"
$NCU $D -k regex:"conv_planes_kernel<\(int\)2, \(int\)4, \(int\)1, \(bool\)0, \(bool\)1>" --launch-skip 9 --launch-count 1 -o $O/r03_conv_c128k11_c1 -f python tools/ncu_decode.py 2 > $O/r03_ncu_a.log 2>&1
$NCU $D -k regex:"rel_attention_umma" --launch-skip 6 --launch-count 1 -o $O/r03_attn_umma_final -f python tools/ncu_decode.py 2 > $O/r03_ncu_b.log 2>&1
$NCU $D -k regex:"conv_planes_kernel<\(int\)1, \(int\)4, \(int\)4" --launch-skip 5 --launch-count 1 -o $O/r03_ups1_fused -f python tools/ncu_decode.py 2 > $O/r03_ncu_c.log 2>&1
for f in r03_conv_c128k11_c1 r03_attn_umma_final r03_ups1_fused; do ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null; done
DIAG_FRAMES=130,97 timeout 600 compute-sanitizer --tool memcheck python tools/diag_det.py 0 v2-48k 3 > $O/r03_sanitizer_memcheck.txt 2>&1
tail -3 $O/r03_sanitizer_memcheck.txt
timeout 300 python tools/diag_det.py 0 v2-48k 400 > $O/r03_determinism.txt 2>&1
tail -2 $O/r03_determinism.txt
for c in v2-32k v2-40k v1-40k; do python bench.py --config $c --no-dropin > $O/r03_bench_$c.json 2>/dev/null; cut -c1-200 $O/r03_bench_$c.json; done
python bench.py --workload batch512 --no-cpu > $O/r03_bench_batch512.json 2>/dev/null; cut -c1-200 $O/r03_bench_batch512.json
