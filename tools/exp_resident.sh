# A/B of the file-resident sweep on one box (under gpurun): lock step as the reference, then the resident
# kernel on the 100 h corpus and on one 60-minute file with its phase breakdown.  VARIANTS: space-separated
# environment settings to compare ("-" = defaults).
echo "== c5 lockstep"; timeout 150 python bench.py --workload c5 --sweep_mode lockstep --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 ms', round(d['ms_per_step'],2), 'files_done', d['run']['files_done'])"
for v in ${VARIANTS:--}; do
  [ "$v" = "-" ] && v="IPFA_NONE=1"
  echo "== c5 resident $v"
  env $v timeout 100 python bench.py --workload c5 --sweep_mode resident --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 ms', round(d['ms_per_step'],2), 'files_done', d['run']['files_done'], 'cap', d['run']['capacity_T_C_K_rank0'])"
  echo "== chain $v"; env $v IPFA_SWEEP_PHASES=1 timeout 60 python tools/check_resident.py 60 1 2>&1 | grep "ipfa resident\|identical\|launch alone" | tail -3 | cut -c1-260
done
