# The cluster kernel (k_coarse_tail) for the LAST levels only: level 5 + dense (290 cells), levels 4-5 + dense (1 184 cells) — round 2 had tried it from level 3 (4.7 k cells) down.
mkdir -p gpurun_out
for cfg in "off:" "t300:MOF_MG_TAIL_CELLS=300" "t1200:MOF_MG_TAIL_CELLS=1200" "t300_c4:MOF_MG_TAIL_CELLS=300 MOF_MG_TAIL_CTAS=4" "t1200_c8:MOF_MG_TAIL_CELLS=1200 MOF_MG_TAIL_CTAS=8"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env MOF_SMOOTH_AHEAD=0 MOF_MG_VERBOSE=1 $envs timeout 300 python tests/diag_timing.py 9 6 ) > gpurun_out/r3d_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r3d_l9_$name.log
  echo "== $name"; grep "small levels" gpurun_out/r3d_l9_$name.log | head -2 | cut -c1-160; grep -E "^it[0-9]|rc |rror" gpurun_out/r3d_l9_$name.log | tail -3 | cut -c1-140
done
