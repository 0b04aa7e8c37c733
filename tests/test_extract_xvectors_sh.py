"""Flow test of x-vector-kaldi-tf_b200/extract_xvectors.sh (the GPU counterpart of the reference's
local/tf/extract_xvectors.sh) with stand-ins for the Kaldi utilities it calls (parse_options.sh, utils/split_data.sh,
run.pl, ivector-mean) and for the python interpreter: which jobs are started on which device with which flags, and what
files the three stages leave behind.  No Kaldi, no GPU."""
import os
import stat
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "x-vector-kaldi-tf_b200", "extract_xvectors.sh")

STUBS = {
    # Kaldi's utils/parse_options.sh, reduced to what the script uses: --some-name value -> some_name=value
    "parse_options.sh": r'''
while [ $# -gt 0 ]; do
  case "$1" in
    --*) _n=$(echo "$1" | sed 's/^--//; s/-/_/g'); eval "$_n=\"\$2\""; shift 2 ;;
    *) break ;;
  esac
done
true
''',
    "run.pl": r'''#!/bin/bash
log=$1; shift
"$@" > "$log" 2>&1
''',
    "ivector-mean": r'''#!/bin/bash
echo "ivector-mean $*"
ark=$(echo "$3" | sed 's/^ark,scp://; s/,.*//'); scp=$(echo "$3" | sed 's/.*,//')
echo "spk1 [ 0 ]" > "$ark"; echo "spk1 $ark:5" > "$scp"; echo "spk1 2" > "${4#ark,t:}"
''',
    # stands in for the interpreter: records how extract_embedding.py was called and writes the job's scp
    "python": r'''#!/bin/bash
echo "XVEC_DEVICE=${XVEC_DEVICE} $*" >> "$STUB_CALLS"
for a in "$@"; do
  case "$a" in
    --vector-wspecifier=*) out=$(echo "$a" | sed 's/.*ark,scp://'); ark=${out%%,*}; scp=${out##*,}
       job=$(basename "$scp" .scp | sed 's/xvector\.//'); echo "data" > "$ark"; echo "utt$job $ark:4" > "$scp" ;;
  esac
done
''',
}


@pytest.fixture()
def recipe(tmp_path):
    bindir = tmp_path / "bin"
    bindir.mkdir()
    (tmp_path / "utils").mkdir()                     # recipes run from the recipe directory, which has utils/ (as the reference assumes)
    for name, body in STUBS.items():
        p = bindir / name
        p.write_text(body)
        p.chmod(p.stat().st_mode | stat.S_IEXEC)
    split = tmp_path / "utils" / "split_data.sh"
    split.write_text(r'''#!/bin/bash
[ "$1" == "--per-utt" ] && shift
data=$1; nj=$2
for j in $(seq $nj); do mkdir -p $data/split${nj}utt/$j; echo "utt$j f.ark:1" > $data/split${nj}utt/$j/feats.scp; echo "utt$j v.ark:1" > $data/split${nj}utt/$j/vad.scp; done
''')
    split.chmod(split.stat().st_mode | stat.S_IEXEC)
    nnet, data = tmp_path / "nnet", tmp_path / "data"
    (nnet / "model_final").mkdir(parents=True)
    (nnet / "model_final" / "model.meta").write_text("{}")
    (nnet / "min_chunk_size").write_text("25\n")
    (nnet / "max_chunk_size").write_text("10000\n")
    data.mkdir()
    for f in ("feats.scp", "vad.scp", "spk2utt"):
        (data / f).write_text("x y\n")
    env = dict(os.environ, PATH="%s:%s" % (bindir, os.environ["PATH"]), STUB_CALLS=str(tmp_path / "calls.txt"))
    return tmp_path, nnet, data, env


def _run(recipe, *options):
    tmp_path, nnet, data, env = recipe
    out = tmp_path / "xvectors"
    r = subprocess.run(["bash", SCRIPT, *options, str(nnet), str(data), str(out)], cwd=str(tmp_path), env=env,
                       capture_output=True, text=True, timeout=120)
    calls = (tmp_path / "calls.txt").read_text().splitlines() if (tmp_path / "calls.txt").exists() else []
    return r, out, calls


def test_three_jobs_on_two_gpus_with_the_device_front_end(recipe):
    r, out, calls = _run(recipe, "--nj", "3", "--num-gpus", "2")
    assert r.returncode == 0, r.stdout + r.stderr
    assert len(calls) == 3
    assert sorted(c.split()[0] for c in calls) == ["XVEC_DEVICE=0", "XVEC_DEVICE=0", "XVEC_DEVICE=1"]   # job j -> GPU (j-1) mod 2
    for c in calls:
        assert "extract_embedding.py" in c and "--use-gpu=yes" in c and "--min-chunk-size=25" in c and "--chunk-size=10000" in c
        assert "--apply-cmvn-sliding=yes" in c and "--cmn-window=300" in c and "--norm-vars=false" in c and "--center=true" in c
        assert "--feature-rspecifier=scp:" in c and "/feats.scp" in c and "--vad-rspecifier=scp,s,cs:" in c and "/vad.scp" in c
        assert "apply-cmvn-sliding --norm-vars" not in c                                       # no Kaldi pipe
        assert "--model-dir=" in c and c.rstrip().endswith("model_final")
    assert (out / "xvector.scp").read_text().split() == ["utt1", str(out / "xvector.1.ark") + ":4",
                                                         "utt2", str(out / "xvector.2.ark") + ":4",
                                                         "utt3", str(out / "xvector.3.ark") + ":4"]
    for f in ("spk_xvector.ark", "spk_xvector.scp", "num_utts.ark", "log/extract.1.log", "log/extract.3.log", "log/speaker_mean.log"):
        assert (out / f).exists(), f


def test_kaldi_pipe_mode_and_chunk_size_checks(recipe):
    r, out, calls = _run(recipe, "--nj", "1", "--num-gpus", "1", "--device-frontend", "false", "--chunk-size", "400")
    assert r.returncode == 0, r.stdout + r.stderr
    assert len(calls) == 1 and "--chunk-size=400" in calls[0]
    assert "apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 scp:" in calls[0]      # the reference's pipe
    assert "select-voiced-frames ark:- scp,s,cs:" in calls[0] and "--apply-cmvn-sliding=yes" not in calls[0]
    r, _, _ = _run(recipe, "--chunk-size", "20000")
    assert r.returncode != 0 and "larger than the maximum chunk size" in (r.stdout + r.stderr)
    r, _, _ = _run(recipe)                                            # usage when the positional arguments are missing
    tmp_path, nnet, data, env = recipe
    r = subprocess.run(["bash", SCRIPT, str(nnet)], cwd=str(tmp_path), env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "Usage:" in r.stderr
