#!/bin/bash
# dev-time GPU call for the slice-coder work: parity (slice_data byte-identical vs the oracle) of the default build, the coder
# stage's time for every wrenc_b200/lib*.so variant (tools/coder_bench.py), and the per-kernel times of one coder pass (ncu launch list)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${TAG:-coder}
SEL='golden or cif_two or qp_sweep or random_and_flat or single_ctu or streaming or extreme or second_walk or arena or decoded_gpu or pinned'
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_reference_clips.py -m gpu -x -q -k "$SEL or clip" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest default rc $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
rm -f gpurun_out/${TAG}_bench.log
for lib in wrenc_b200/libwrenc_b200.so wrenc_b200/lib_*.so; do
  [ -f "$lib" ] || continue
  WRENC_B200_LIB=$GRAFT_REPO_ROOT/$lib timeout 300 python tools/coder_bench.py >> gpurun_out/${TAG}_bench.log 2>&1
done
cat gpurun_out/${TAG}_bench.log
FS=240 REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/coder_bench.py > gpurun_out/${TAG}_ncu.log 2>&1
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/'+__import__('os').environ.get('TAG','coder')+'_launches.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]; hd=rows[h]
for r in rows[h+2:]:
    if len(r)>=len(hd) and 'wrenc' in r[hd.index('Kernel Name')]: print(r[hd.index('Kernel Name')][:40], r[hd.index('Metric Value')])
P
