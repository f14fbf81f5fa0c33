"""Admission control of the contigs in flight on one GPU (grom_b200/pipeline.py: _InFlight) -- host logic, no GPU."""
import threading
import time

from grom_b200.pipeline import _InFlight


def test_budget_is_respected_and_an_oversize_contig_runs_alone():
    fl = _InFlight(budget=100)
    log, lock = [], threading.Lock()
    peak = [0]

    def job(need, hold):
        fl.acquire(need)
        with lock:
            log.append(("in", need, fl.used, fl.running))
            peak[0] = max(peak[0], fl.used if fl.running > 1 else 0)
        time.sleep(hold)
        fl.release(need)

    th = [threading.Thread(target=job, args=(n, 0.05)) for n in (60, 30, 30, 250, 10, 40)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert fl.used == 0 and fl.running == 0 and len(log) == 6
    assert peak[0] <= 100                                     # several in flight never exceed the budget
    big = [e for e in log if e[1] == 250][0]
    assert big[3] == 1 and big[2] == 250                      # what fits nowhere is admitted only when nothing else runs
