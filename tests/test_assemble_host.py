"""Host-side pieces of bear_b200.assemble (reference assemble.py): no GPU needed."""
import numpy as np
import pytest


def test_reverse_complement_and_fasta(tmp_path):
    from bear_b200 import assemble
    assert assemble.reverse_complement('AACGT') == 'ACGTT'
    assert assemble.reverse_complement('AACGU', 'rna') == 'ACGUU'
    fa = tmp_path / 's.fa'
    fa.write_text('>a desc\nACG\nTT\n\n>b\nGG\n')
    assert assemble.read_fasta(str(fa)) == ['ACGTT', 'GG']


def test_sitewise_entropy():
    from bear_b200 import assemble
    ent = assemble.sitewise_entropy(['ACG', 'ACT', 'AGT', 'ATT'], 'dna')
    want = [0.0, -(0.5 * np.log(0.5) + 2 * 0.25 * np.log(0.25)), -(0.25 * np.log(0.25) + 0.75 * np.log(0.75))]
    assert np.allclose(ent, want)


def test_kmc_counter_is_refused():
    from bear_b200 import assemble
    with pytest.raises(NotImplementedError):
        assemble.assemble_no_ends(['ACGT'], [[0, 1]], 1, None, kmc_path='db.res', van=1.0, lag=2, alphabet_name='dna')
