set -x
python -m pytest tests -m gpu -x -q -s > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?" >> gpurun_out/s3_bench.err
timeout 200 python profiles/backbones_bench.py > gpurun_out/s3_backbones.txt 2>&1
timeout 120 python profiles/prof_backbone.py clip_vit_l_14 146 >> gpurun_out/s3_backbones.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s3_launches.csv python profiles/prof_step.py 3 > gpurun_out/s3_ncu_launch.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attention_ppl -c 1 -o gpurun_out/s3_attn_ppl python profiles/attn_long_bench.py 96 > gpurun_out/s3_ncu_full.log 2>&1
timeout 60 python profiles/ncu_kernel_summary.py gpurun_out/s3_attn_ppl.ncu-rep > gpurun_out/s3_attn_ppl_summary.txt 2>&1
tail -3 gpurun_out/s3_pytest.log; tail -2 gpurun_out/s3_bench.err; tail -4 gpurun_out/s3_backbones.txt
