python -m pytest tests -m gpu -q 2>&1 | tail -3
for t in 1 2 3 4; do
  TTM_FIT_THREADS=$t python tools/fit_c4.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads', d['optimize_s'], d['fused_evals_this_rank'], d['max_abs_grad_at_solution'], d['sum_J'])"
done
