export B2S_REFERENCE_DIR=$PWD/baseline/_ref/code
run() { python tests/e2e_prove_dropin.py gpu gpurun_out/g_$3.json "$1" "$2" $3.json > /dev/null 2>&1; python -c "
import json; d=json.load(open('gpurun_out/g_$3.json')); print('$3', d['fri_domain_length'], 'prove', d['prove_seconds'], 's  reference', d.get('reference_prove_seconds'), 's  byte-identical', d.get('byte_identical_to_reference_proof'), ' verifier', d['reference_verifier_accepts'])"; }
run "++++" "" bfs
run "++[>,.<-]" "ab" bfs_io
run "+++++[>,.<-]" "hello" bfs_echo
run "+++[>+++[>+<-]<-]>>." "" bfs_nested
run "++++++++[>++++++++<-]>+." "" bfs_A
run "++++++++++[>+++++++>++++++++++<<-]>++.>+." "" bfs_He
python - <<'PY'
import json, glob
out = {}
for f in sorted(glob.glob('gpurun_out/g_bfs*.json')):
    d = json.load(open(f)); out[f.split('g_')[1][:-5]] = {k: d[k] for k in ('program','fri_domain_length','prove_seconds','reference_prove_seconds','byte_identical_to_reference_proof','reference_verifier_accepts','proof_sha256') if k in d}
json.dump(out, open('gpurun_out/r02cg_all_golden_programs_b200.json','w'), indent=1)
PY
