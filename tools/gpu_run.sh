#!/bin/bash
# One gpurun call = several measurements, each bounded by its own timeout, logs under gpurun_out/<tag>_*.log.
# usage: tools/gpu_run.sh <tag> <step> [<step> ...]   steps: smoke tests tests_fast bench_c3 bench_c4 bench_c5 ncu_list ncu_chol
tag=$1; shift
out=gpurun_out; mkdir -p $out
for step in "$@"; do
  echo "=== $step"; t0=$(date +%s)
  case $step in
    smoke)      timeout 600 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "rc=$?"; tail -5 $out/${tag}_smoke.log ;;
    tests)      timeout 3000 python -m pytest tests -m gpu -q --timeout 1500 -p no:cacheprovider > $out/${tag}_tests.log 2>&1; echo "rc=$?"; tail -40 $out/${tag}_tests.log ;;
    tests_fast) timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --deselect tests/test_gpu_fullsize_c4c5.py --deselect tests/test_gpu_fullsize.py > $out/${tag}_tests.log 2>&1; echo "rc=$?"; tail -40 $out/${tag}_tests.log ;;
    tests_multi) timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider > $out/${tag}_tests_multi.log 2>&1; echo "rc=$?"; tail -30 $out/${tag}_tests_multi.log ;;
    tests_full) timeout 2400 python -m pytest tests/test_gpu_fullsize_c4c5.py tests/test_gpu_fullsize.py -m gpu -q --timeout 1500 -p no:cacheprovider > $out/${tag}_tests_full.log 2>&1; echo "rc=$?"; tail -30 $out/${tag}_tests_full.log ;;
    bench_c3)   timeout 900 python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err; echo "rc=$?"; tail -c 600 $out/${tag}_bench_c3.err; head -c 1500 $out/${tag}_bench_c3.json ;;
    bench_c3_fast) timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-setup > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err; echo "rc=$?"; tail -c 600 $out/${tag}_bench_c3.err; head -c 1500 $out/${tag}_bench_c3.json ;;
    bench_c4)   timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err; echo "rc=$?"; tail -c 600 $out/${tag}_bench_c4.err; head -c 1200 $out/${tag}_bench_c4.json ;;
    bench_c5)   timeout 900 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; echo "rc=$?"; tail -c 600 $out/${tag}_bench_c5.err; head -c 1200 $out/${tag}_bench_c5.json ;;
    bench_c3_n*) n=${step#bench_c3_n}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > $out/${tag}_bench_c3_n$n.json 2> $out/${tag}_bench_c3_n$n.err; echo "rc=$?"; tail -c 1500 $out/${tag}_bench_c3_n$n.err; head -c 700 $out/${tag}_bench_c3_n$n.json ;;
    bench_c4_n*) n=${step#bench_c4_n}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --workload c4 --steps 2 --warmup 1 --no-setup > $out/${tag}_bench_c4_n$n.json 2> $out/${tag}_bench_c4_n$n.err; echo "rc=$?"; tail -c 1500 $out/${tag}_bench_c4_n$n.err; head -c 700 $out/${tag}_bench_c4_n$n.json ;;
    bench_c5_n*) n=${step#bench_c5_n}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --workload c5 --steps 2 --warmup 1 --no-setup > $out/${tag}_bench_c5_n$n.json 2> $out/${tag}_bench_c5_n$n.err; echo "rc=$?"; tail -c 1500 $out/${tag}_bench_c5_n$n.err; head -c 700 $out/${tag}_bench_c5_n$n.json ;;
    ref_n*)     n=${step#ref_n}; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $n --steps 20 --warmup 3 > $out/${tag}_ref_n$n.json 2> $out/${tag}_ref_n$n.err; echo "rc=$?"; tail -c 600 $out/${tag}_ref_n$n.err; head -c 900 $out/${tag}_ref_n$n.json ;;
    tests_failed) timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "checkpoint or gram_stationary or alpha_refinement or dense_debug" > $out/${tag}_tests_failed.log 2>&1; echo "rc=$?"; tail -15 $out/${tag}_tests_failed.log ;;
    fit_ab)     for v in -1 0 1; do for w in c3 c4 c3_8th; do echo -n "PMK_CHOL_VARIANT=$v $w: "; PMK_CHOL_VARIANT=$v timeout 200 python tools/fit_only.py $w 3 2>&1 | tail -1; done; done ;;
    tests_fit_levels) PMK_CHOL_VARIANT=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider -k "fit or leaf_size or positive_definite or large_leaf or golden or multi or ibb1d or checkpoint or alpha or cholesky" > $out/${tag}_tests_fit_levels.log 2>&1; echo "rc=$?"; tail -15 $out/${tag}_tests_fit_levels.log ;;
    tests_fit)  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_multi.py -m gpu -q --timeout 600 -p no:cacheprovider -k "fit or leaf_size or positive_definite or large_leaf or golden or multi or ibb1d or checkpoint or alpha or cholesky" > $out/${tag}_tests_fit.log 2>&1; echo "rc=$?"; tail -15 $out/${tag}_tests_fit.log ;;
    ref_arm)    timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "rc=$?"; head -c 800 $out/${tag}_bench_ref.json ;;
    ncu_list)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-setup > $out/${tag}_ncu_list.log 2>&1; echo "rc=$?" ;;
    ncu_k3)     timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_query_rowp' -c 2 -o $out/${tag}_k3 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-setup > $out/${tag}_ncu_k3.log 2>&1; echo "rc=$?"; tail -3 $out/${tag}_ncu_k3.log; ncu -i $out/${tag}_k3.ncu-rep --page raw --csv > $out/${tag}_k3_raw.csv 2>/dev/null; ls -la $out/${tag}_k3* ;;
    ncu_chol)   timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_chol|k_solve_alpha|k_gram_tiles' -c 3 -o $out/${tag}_chol python tools/fit_only.py c3 1 > $out/${tag}_ncu_chol.log 2>&1; echo "rc=$?"; tail -3 $out/${tag}_ncu_chol.log; ncu -i $out/${tag}_chol.ncu-rep --page raw --csv > $out/${tag}_chol_raw.csv 2>/dev/null; ls -la $out/${tag}_chol* ;;
    ncu_fit)    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_chol|k_gram_tiles|k_solve_alpha|k_inv_' -c 30 -o $out/${tag}_fit python tools/fit_only.py > $out/${tag}_ncu_fit.log 2>&1; echo "rc=$?"; tail -3 $out/${tag}_ncu_fit.log ;;
    *) echo "unknown step $step" ;;
  esac
  echo "--- $step took $(( $(date +%s) - t0 )) s"
done
nvidia-smi --query-gpu=name,memory.used,clocks.sm --format=csv,noheader
