mkdir -p gpurun_out
echo "== C2 kernel variants (probe: 4096 chains x 1024 particles, 1024 events per step)"
for lib in build_variants/base.so build_variants/a_unroll2.so build_variants/b_prefilter.so build_variants/ab.so build_variants/base.so; do
  echo "-- $lib"
  JELLYFYSH_B200_LIBRARY=$PWD/$lib timeout 200 python tools/probe.py 4096 1024 12 1024 2>&1 | grep -E "step [2-5]|rror" | cut -c1-150
done
echo "== timeline of pipelined host steps (staged full copy)"
timeout 200 python tools/probe_timeline.py 0 3 > gpurun_out/r2t_timeline_full.txt 2>&1; head -14 gpurun_out/r2t_timeline_full.txt | cut -c1-150
echo "== timeline (sparse, fused)"
timeout 200 python tools/probe_timeline.py 1 3 > gpurun_out/r2t_timeline_sparse.txt 2>&1; head -10 gpurun_out/r2t_timeline_sparse.txt | cut -c1-150
