run() { echo "== $1"; env $1 python tools/quick_perf.py 0 20,22,24 2>&1 | grep log_n | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['log_n'],'reduce',d['reduce'],'total',d['total'])"; }
run "X=0"
run "MIRA_RED_COOP_MAX_LOG=19"
run "MIRA_RED1_MIN_LOG=19"
run "MIRA_RED1_LOG_M=3"
run "MIRA_RED1_LOG_M=5"
run "MIRA_RED1_NEXT_MIN_LOG=11"
run "MIRA_RED1_NEXT_MIN_LOG=15"
run "MIRA_RED1_NEXT_MIN_LOG=17"
