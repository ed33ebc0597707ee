"""Table and rule check for a survey written by tools/k2grid_survey.py (no GPU needed: the plan quantities come from the
host planners).  usage: python tools/k2grid_survey_table.py profiles/r02_k2_grid_survey.jsonl
Prints, per grid, the fraction of the FP64 peak of every kernel, the fastest one, the plan quantities
(kb2_self_moments_tile_plan_cost, kb2_self_moments_plan, kb2_self_moments_win_plan) and what kb2_self_moments_choose takes;
then the geometric mean and the worst case of chosen / fastest time, and the same for always taking one kernel."""
import json
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from kinisi_b200 import _lib, engine

PEAK = 35.5e12  # FP64 flop/s (DFMA micro-benchmark of bench.py)
lib = _lib.load()
rows = []
print('     T n_st used atoms | frac0 frac2 frac3 frac4 | best chosen | reload B/term cyc/t | runs(n,len,chg,rest) m_win | err')
for line in open(sys.argv[1]):
    r = json.loads(line)
    T, n = r['T'], r['n_steps']
    lags = np.unique(np.linspace(1, T, n, dtype=int))[:r['n_used']].astype(np.int32)
    cost = np.zeros(3)
    assert lib.kb2_self_moments_tile_plan_cost(lags.ctypes.data, len(lags), T, cost.ctypes.data) == 0
    terms = float(np.sum(T - lags + 1))
    runs, rest = engine.self_moments_plan(lags[:192], T)
    wins = engine.self_moments_windows(lags[:256], T)
    frac = {int(v): 16 * r['terms'] / (ms * 1e-3) / PEAK for v, ms in r['ms'].items()}
    best = max(frac, key=frac.get)
    chosen = engine.choose_variant(lags, T)
    rows.append((frac, best, chosen))
    run_len = np.mean([x[2] for x in runs]) if runs else 0
    print(f"{T:6d} {n:4d} {r['n_used']:4d} {r['n_atoms']:5d} | {frac[0]:.2f}  {frac[2]:.2f}  {frac[3]:.2f}  {frac[4]:.2f} |  {best}    {chosen}    |"
          f" {cost[0] / (24 * T):6.2f} {cost[0] / terms:6.2f} {cost[1] / terms * 32:5.1f} |"
          f" ({len(runs)},{run_len:.0f},{sum(abs(x[3]) for x in runs)},{rest}) {np.mean([w[2] for w in wins]):.2f} | {max(r['err'].values()):.0e}")


def regret(pick):
    ratios = [max(f.values()) / f[pick(f, c)] for f, _b, c in rows]
    return float(np.exp(np.mean(np.log(ratios)))), float(max(ratios))


print('chosen / fastest time: geometric mean %.3f, worst %.2f' % regret(lambda f, c: c))
for v, name in ((4, 'tiles'), (3, 'windows'), (2, 'runs'), (0, 'direct')):
    print('always %-8s: geometric mean %.3f, worst %.2f' % ((name,) + regret(lambda f, c, v=v: v)))
