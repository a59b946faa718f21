"""A/B driver for decode-kernel build variants (how the table in profiles/r01_decode.md was measured).

Build a variant next to the default library, e.g.
    CERB_OUT=cerberusdet_b200/libcerb_b4.so sh cerberusdet_b200/csrc/build.sh -DDEC_MINB=4
    CERB_OUT=cerberusdet_b200/libcerb_db.so sh cerberusdet_b200/csrc/build.sh -DDEC_DOUBLE_BUFFER -DDEC_MINB=3
then on the GPU box
    python tools/decode_variants.py default b4=libcerb_b4.so db=libcerb_db.so tma=:CERB_DEBUG_DECODE_TMA=1 order1=:CERB_DEBUG_DECODE_ORDER=1
Each spec is name[=lib][:ENV=VALUE,...]; every variant runs tools/microbench.py cfg3 cfg3f32 in its own process.
Compile-time switches: DEC_THREADS, DEC_MINB, DEC_DOUBLE_BUFFER, CERB_L2_PREFETCH, CERB_LD_L2_128/256, CERB_LD_PLAIN,
CERB_EXPERIMENT_NO_MUFU, CERB_EXPERIMENT_COPY_ONLY (the last two give wrong results on purpose).
"""
import json, os, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in sys.argv[1:] or ["default"]:
    head, _, envs = spec.partition(":")
    name, _, lib = head.partition("=")
    env = dict(os.environ)
    if lib:
        env["CERB_LIB"] = os.path.join(root, "cerberusdet_b200", lib)
    for kv in filter(None, envs.split(",")):
        k, _, v = kv.partition("=")
        env[k] = v
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "microbench.py"), "cfg3", "cfg3f32"], env=env,
                         capture_output=True, text=True).stdout.strip().splitlines()
    for line in out[-2:]:
        d = json.loads(line)
        print(name, d["cfg"], d["decode_us_med"], d["decode_us_min"], d["decode_GBps_med"], flush=True)
