# memcheck over the small / medium GPU tests that exercise the kernels written or rewritten in round 2
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 30 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
  -k "s2n_bit_exact or perm_null_s2n_wks or mean_maxmean and not c4 or tail_round or near_duplicated or few_samples or user_score_tiles or calc_sig_counts or staged_sig or two_responses or small_n_label or observed_es_is_copied or custom_es_fn or tiny_universe or inf_patch or lm_scores_and_null" \
  > gpurun_out/r2_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_memcheck.log
tail -25 gpurun_out/r2_memcheck.log
