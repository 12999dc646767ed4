python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/final_tests.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-others > gpurun_out/final_launch_bench.log 2>&1
for w in "tweets 1000000 3 tweets" "tweets 1000000 7 tweets_feats" "mixed 1000000 3 mixed" "mixed 1000000 7 mixed_feats" "docs 3000 3 docs"; do set -- $w; ncu --set full --clock-control none --import-source on -k regex:tokenize5 -s 2 -c 1 -f -o gpurun_out/final_$4 python tools/prof_run.py $1 $2 4 $3 > gpurun_out/final_ncu_$4.log 2>&1; done
cat gpurun_out/final_tests.log; head -c 1500 gpurun_out/final_bench.json; ls -la gpurun_out | tail -12
