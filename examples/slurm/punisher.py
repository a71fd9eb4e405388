"""Chaos tool: kill replica groups of a running job to exercise fault tolerance.

    python examples/slurm/punisher.py kill_one  --lighthouse http://host:29510
    python examples/slurm/punisher.py kill_all  --lighthouse http://host:29510
    python examples/slurm/punisher.py kill_loop --lighthouse http://host:29510 --mtbf-secs 300

Kills go through the Lighthouse dashboard endpoint ``POST /replica/{id}/kill`` (which forwards a
``Kill`` RPC to that replica's ManagerServer -> ``exit(1)``), so this works on any scheduler.
``kill_loop`` draws exponential inter-failure times with the requested mean (the reference's
version computes such a draw and then sleeps the fixed MTBF, punisher.py:49-51).
"""

from __future__ import annotations

import argparse
import json
import random
import re
import time
import urllib.error
import urllib.parse
import urllib.request
from typing import List


def replicas(lighthouse: str) -> List[str]:
    """Replica ids of the current quorum (GET /status.json; falls back to scraping the HTML dashboard)."""
    base = lighthouse.rstrip("/")
    try:
        st = json.loads(urllib.request.urlopen(base + "/status.json", timeout=10).read().decode())
        return [p["replica_id"] for p in (st.get("prev_quorum") or {}).get("participants", [])]
    except (urllib.error.HTTPError, ValueError):
        html = urllib.request.urlopen(base + "/status", timeout=10).read().decode()
        return re.findall(r"kill\('([^']+)'\)", html)


def kill(lighthouse: str, replica_id: str) -> None:
    url = f"{lighthouse.rstrip('/')}/replica/{urllib.parse.quote(replica_id, safe='')}/kill"
    try:
        urllib.request.urlopen(urllib.request.Request(url, method="POST"), timeout=15).read()
    except Exception as e:  # the target dies before it can answer
        print(f"kill {replica_id}: {e}")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["kill_one", "kill_all", "kill_loop"])
    ap.add_argument("--lighthouse", required=True)
    ap.add_argument("--mtbf-secs", type=float, default=300.0)
    a = ap.parse_args()
    if a.cmd == "kill_one":
        ids = replicas(a.lighthouse)
        if ids:
            kill(a.lighthouse, random.choice(ids))
    elif a.cmd == "kill_all":
        for rid in replicas(a.lighthouse):
            kill(a.lighthouse, rid)
    else:
        while True:
            time.sleep(random.expovariate(1.0 / a.mtbf_secs))
            ids = replicas(a.lighthouse)
            if ids:
                victim = random.choice(ids)
                print(f"killing {victim}")
                kill(a.lighthouse, victim)


if __name__ == "__main__":
    main()
