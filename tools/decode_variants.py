import os, sys, subprocess, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = [("default", None, None)] + [(f"t{t}_b{b}", f"{root}/cerberusdet_b200/libcerb_dec_{t}_{b}.so", None) for t, b in [(128,4),(128,5),(256,2),(256,3),(64,8)]]
variants += [("default_vec4", None, "4"), ("t128_b5_vec4", f"{root}/cerberusdet_b200/libcerb_dec_128_5.so", "4"), ("t256_b3_vec4", f"{root}/cerberusdet_b200/libcerb_dec_256_3.so", "4")]
for name, lib, vec in variants:
    env = dict(os.environ)
    if lib: env["CERB_LIB"] = lib
    if vec: env["CERB_DEBUG_DECODE_VEC"] = vec
    out = subprocess.run([sys.executable, f"{root}/tools/microbench.py", "cfg3"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
    d = json.loads(out[-1]) if out else {}
    print(name, d.get("decode_us_med"), d.get("decode_us_min"), d.get("decode_GBps_med"), flush=True)
