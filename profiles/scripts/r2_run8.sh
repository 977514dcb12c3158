set -x
mkdir -p gpurun_out
: > gpurun_out/r2_es_variants.jsonl
python profiles/es_variants.py >> gpurun_out/r2_es_variants.jsonl 2> gpurun_out/r2_es_variants.err
for v in xorcmp rankbank both; do
  FLEXGSEA_B200_LIB=$PWD/profiles/variants/libflexgsea_b200_$v.so python profiles/es_variants.py >> gpurun_out/r2_es_variants.jsonl 2>> gpurun_out/r2_es_variants.err
done
python profiles/es_variants.py >> gpurun_out/r2_es_variants.jsonl 2>> gpurun_out/r2_es_variants.err
cat gpurun_out/r2_es_variants.jsonl
