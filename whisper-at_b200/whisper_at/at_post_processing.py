"""Label post-processing with the reference's names, arguments and output structure
(package/whisper-at/whisper_at/at_post_processing.py:14-62).  Label names come from
assets/at_labels.json.gz (built by tools/make_assets.py from the reference's data assets)."""
from __future__ import annotations

import gzip
import json
import os
import warnings
from functools import lru_cache

import torch


@lru_cache(maxsize=1)
def _assets():
    path = os.path.join(os.path.dirname(__file__), "assets", "at_labels.json.gz")
    with gzip.open(path, "rt", encoding="utf-8") as f:
        return json.load(f)


def parse_at_label(result, language='follow_asr', top_k=5, p_threshold=-1, include_class_list=list(range(527))):
    """Top-k tags per `at_time_res` window.

    result: dict returned by transcribe(); language: label language ('follow_asr' = result['language']);
    keeps the top_k classes whose logit exceeds p_threshold and whose index is in include_class_list.
    Returns [{'time': {'start', 'end'}, 'audio tags': [(name, logit), ...]}, ...]."""
    labels_by_lang = _assets()["labels"]
    res = result['at_time_res']
    tags = result['audio_tag']
    lang = result['language'] if language == 'follow_asr' else language
    if lang not in labels_by_lang:
        warnings.warn("{:s} language not supported. Use English label names instead. If you wish to use label names of a specific language, please specify the language argument".format(lang))
        lang = 'en'
    names = labels_by_lang[lang]
    out = []
    for w in range(tags.shape[0]):
        values, indices = torch.topk(tags[w], k=top_k)
        picked = []
        for v, i in zip(values, indices):
            if v > p_threshold and i in include_class_list:
                picked.append((names[i], v.item()))
        out.append({'time': {'start': w * res, 'end': (w + 1) * res}, 'audio tags': picked})
    return out


def print_label_name(language='en'):
    for i, name in enumerate(_assets()["labels"][language]):
        print("index: {:d} : {:s}".format(i, name))


def print_support_language():
    a = _assets()
    for key in a["labels"].keys():
        print("language code: {:s} : {:s}".format(key, a["language_names"][key]))
