"""print the e2e block of a bench line: python scripts/show_e2e.py bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d["configs"]["e2e"]
print({k: e[k] for k in ("db_build_s", "db_subsessions_per_s", "db_encode_share", "query_sessions_per_s_end_to_end",
                         "query_path_s", "query_path_share", "parity")})
print(e["encoder"])
