#!/bin/bash
# Drop-in shadows for the reference's back-end shell functions.  Source this file AFTER the reference's own
# function files, i.e. in scoreSets.sh after lines 133-134:
#
#     . subtools/score/process.sh
#     . subtools/score/score.sh
#     . /path/to/repo/integration/score_b200.sh        # <- the one added line
#
# Every function below keeps the name, the positionals and the output files of the function it shadows
# (reference file:line in the comment above it), so get_params / process / scoreSets.sh keep working unchanged; the
# Kaldi binary (or Python script) behind each step is replaced by the B200 back end
# (python -m asv_subtools_b200.score.*), one (N, D) device tensor per step instead of an ark pipe per vector.
#
#   XVB200_ROOT    repo root (default: the directory above this file)
#   XVB200_PYTHON  interpreter (default: python3)

XVB200_ROOT=${XVB200_ROOT:-$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)}
XVB200_PYTHON=${XVB200_PYTHON:-python3}

function _xvb200(){
	PYTHONPATH="$XVB200_ROOT${PYTHONPATH:+:$PYTHONPATH}" $XVB200_PYTHON -m "asv_subtools_b200.score.$1" "${@:2}"
}

# score/score.sh:82-97 -- ivector-compute-dot-products over the listed trials
function cosine(){
	_xvb200 cosine "$1" "$2" "$3" "$4" || exit 1
	return 0
}

# score/score.sh:99-121 -- ivector-plda-scoring --normalize-length=true [--num-utts] "ivector-copy-plda --smoothing"
function plda(){
	local num_utt=""
	[ -f "$2" ] && num_utt=$2
	_xvb200 plda --kaldi --smoothing "${plda_smoothing:-0.0}" --normalize-length true "$1" "$num_utt" "$3" "$4" "$5" "$6" || exit 1
	return 0
}

# score/score.sh:123-145 -- the same scorer on an adapted model
function aplda(){
	local num_utt=""
	[ -f "$2" ] && num_utt=$2
	_xvb200 plda --kaldi --smoothing "${aplda_smoothing:-0.0}" --normalize-length true "$1" "$num_utt" "$3" "$4" "$5" "$6" || exit 1
	return 0
}

# score/process.sh:156-167 -- ivector-mean ark:spk2utt ... ark:out ark,t:num_utts
function mean(){
	_xvb200 process mean "$1" "$2" "$3" "$4" || exit 1
	return 0
}

# score/process.sh:169-179 -- ivector-mean <vectors> <mean.vec>
function getmean(){
	_xvb200 process getmean "$1" "$2" || exit 1
	return 0
}

# score/process.sh:181-192 -- ivector-subtract-global-mean
function submean(){
	_xvb200 process submean "$1" "$2" "$3" || exit 1
	return 0
}

# score/process.sh:194-203 -- ivector-normalize-length --scaleup=false
function norm(){
	_xvb200 process norm "$1" "$2" || exit 1
	return 0
}

# score/process.sh:205-216 -- ivector-transform (lda() and whiten() of :230-233 / :262-265 call this one)
function transform(){
	_xvb200 process transform "$1" "$2" "$3" || exit 1
	return 0
}

# score/process.sh:218-228 -- ivector-compute-lda --dim=$clda --total-covariance-factor=0.1
function trainlda(){
	_xvb200 process trainlda --dim "${clda:-10}" --total-covariance-factor 0.1 "$1" "$2" "$3" || exit 1
	return 0
}

# score/process.sh:235-248 -- copy-vector to text + score/whiten/train_ZCA_Whitening.py
function trainwhiten(){
	_xvb200 process trainwhiten "$1" "$2" 1>&2 || exit 1
	return 0
}

# score/process.sh:250-260 -- est-pca --read-vectors=true
function trainpcawhiten(){
	_xvb200 process trainpcawhiten "$1" "$2" || exit 1
	return 0
}

# score/process.sh:267-278 -- ivector-compute-plda ark:spk2utt <vectors> <plda>  (EM on the GPU; also writes <plda>.ori)
function trainplda(){
	_xvb200 compute_plda "$1" "$2" "$3" || exit 1
	return 0
}

# score/process.sh:280-292 -- ivector-adapt-plda (Kaldi's unsupervised adaptation; the reference restates it as
# score/pyplda/plda_base.py PldaUnsupervisedAdaptor, which is what --method kaldi mirrors)
function trainaplda(){
	local src=$1
	[ -f "$1.ori" ] && src=$1.ori
	_xvb200 adapt_plda --method kaldi --within-covar-scale "${within_covar_scale:-0.70}" \
		--between-covar-scale "${between_covar_scale:-0.30}" --mean-diff-scale "${mean_diff_scale:-1}" "$src" "$2" "$3" || exit 1
	return 0
}
