"""bioshell_b200 -- B200-native (sm_100a) all-vs-all global alignment behind
BioShell's aligner API.  See DESIGN.md; the C ABI is include/bioshell_align.h."""
from ._lib import BsaError, LIB_PATH  # noqa: F401
from .alignment import (AlignmentPath, AlignmentReporter, AlignmentStatistics, AlignmentStep, CollectReporter, Context,  # noqa: F401
                        GlobalAligner, LocalAlignment, MultiReporter, PairResults, SequenceIdentityMatrix,
                        align_all_pairs, align_all_vs_all, align_one_vs_many, align_pairs_batched,
                        aligned_sequences, aligned_strings, aligned_symbols, triangle_counts)
from .scoring import SubstitutionMatrix, SubstitutionMatrixList, ncbi_text  # noqa: F401
from .sequence import Sequence, count_identical, len_ungapped, pack, ungapped_lengths  # noqa: F401
from . import clustering  # noqa: F401,E402
from . import bucket_clustering  # noqa: F401,E402
from .fasta import FastaIterator, load_sequences  # noqa: F401,E402
from .sequence_id import LabelStyle, SeqId, SeqIdList, parse_sequence_id, sequence_label  # noqa: F401,E402
from .reporters import (IdentityMatrixReporter, PrintAsFasta, PrintAsPairwise,  # noqa: F401,E402
                        ReportWithSequenceIdentity, SimilarityReport)
