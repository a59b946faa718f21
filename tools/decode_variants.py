import os, sys, subprocess, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = [("default(prefetch,b5)", None), ("no_prefetch_b5", "libcerb_nopf.so"), ("prefetch_b3", "libcerb_pf_b3.so"), ("prefetch_b4", "libcerb_pf_b4.so"), ("prefetch_b6", "libcerb_pf_b6.so")]
for name, lib in variants:
    env = dict(os.environ)
    if lib: env["CERB_LIB"] = f"{root}/cerberusdet_b200/{lib}"
    out = subprocess.run([sys.executable, f"{root}/tools/microbench.py", "cfg3", "cfg3f32"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
    for line in out[-2:]:
        d = json.loads(line)
        print(name, d["cfg"], d.get("decode_us_med"), d.get("decode_us_min"), d.get("decode_GBps_med"), flush=True)
