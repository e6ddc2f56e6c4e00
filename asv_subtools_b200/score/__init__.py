"""Back-end scoring on the B200: drop-in for score/score.sh `cosine`/`plda`, the `norm`/`submean`
steps of score/process.sh and the EER scripts."""
