#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2h_pytest.log
for gm in 0 1; do
  echo "== stream_dev group_major=$gm"
  PMU_GROUP_MAJOR=$gm python tools/stream_dev.py config.ini 8 40
  PMU_GROUP_MAJOR=$gm python tools/stream_dev.py config.ini 8 20 trig
  PMU_GROUP_MAJOR=$gm python tools/stream_dev.py cfg5_600 8 40
  PMU_GROUP_MAJOR=$gm python tools/stream_dev.py m_class_config.ini 8 40
  PMU_GROUP_MAJOR=$gm timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:pmu_fused -s 6 -c 2 --csv --log-file gpurun_out/r2h_ncu_gm$gm.csv python tools/stream_dev.py config.ini 8 6 > /dev/null 2>&1
  grep -a "pmu_fused" gpurun_out/r2h_ncu_gm$gm.csv | cut -d, -f5,13-15 | head -8
done
CASES=c3,c3_1f,c3_trig,c5,mc,mc_trig,c3_f32 bash tools/gpu_r2b.sh main
