"""Build whisper_at/assets/at_labels.json.gz from the reference's DATA assets (not source code):
the AudioSet class names in 85 languages (assets/label_name_dict.json, used by parse_at_label,
at_post_processing.py:28-36) and the language-code -> language-name table printed by
print_support_language (tokenizer.py LANGUAGES).  Run in the build container only:

    python tools/make_assets.py
"""
import gzip
import json
import os
import sys

REF = "/root/reference/package/whisper-at/whisper_at"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "whisper-at_b200", "whisper_at",
                   "assets", "at_labels.json.gz")

sys.dont_write_bytecode = True
sys.path.insert(0, os.path.dirname(REF))
from whisper_at.tokenizer import LANGUAGES  # noqa: E402

with open(os.path.join(REF, "assets", "label_name_dict.json")) as f:
    labels = json.load(f)
assert all(len(v) == 527 for v in labels.values())
payload = {"labels": labels, "language_names": {k: LANGUAGES[k] for k in labels if k in LANGUAGES}}
os.makedirs(os.path.dirname(OUT), exist_ok=True)
with gzip.GzipFile(OUT, "wb", mtime=0) as f:
    f.write(json.dumps(payload, ensure_ascii=False, separators=(",", ":")).encode("utf-8"))
print(OUT, os.path.getsize(OUT), "bytes;", len(labels), "languages")
