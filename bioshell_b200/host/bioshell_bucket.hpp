// bioshell_bucket.hpp -- compiled host driver of bioshell-seq's k-mer bucket clustering (CD-HIT-like
// greedy incremental clustering) over the C ABI.  Same names and decisions as the reference:
//
//   bucket_clustering / bucket_clustering_n     bioshell-seq/src/sequence/bucket_clustering/bucket_clustering.rs:33-67
//   BucketClustering::{new, run, run_n, merge, sequence_identity}          .../bucket_clustering.rs:141-309
//   generate_kmers, count_intersection_sorted, kmer_identity_bounds, suggest_word_length   .../kmers.rs:17-121
//   standard_letter_to_index                    bioshell-seq/src/chemical/residue_types.rs:537-580
//
// The reference walks the representatives one at a time and calls GlobalAligner (BLOSUM62, -11 / -1)
// whenever the k-mer bounds are inconclusive, stopping at the first hit.  Here the k-mer verdicts of a
// candidate against the current representatives are taken first (host), and the inconclusive ones that
// precede the first certain hit go to the GPU as ONE batch -- query = representative (rows), template =
// the candidate (columns), exactly the reference's orientation -- through bsa_gather_sequences +
// bsa_align_all_pairs: the forward score + identity kernels, no direction store and no traceback.  The
// batch is then scanned in the reference's order, so the clustering is identical; only the `aligned`
// statistic can exceed the reference's.  There is no CPU alignment path.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "bioshell_seq.hpp"

namespace bioshell_seq {

constexpr uint8_t kInvalidLetter = 255;

// StandardResidueType::TYPES code1 order, residue_types.rs:499-533; later entries overwrite earlier ones (:540-542)
inline const uint8_t* standard_letter_table() {
    static uint8_t t[256];
    static bool init = false;
    if (!init) {
        std::fill(t, t + 256, kInvalidLetter);
        const char* order = "ARNDCQEGHILKMFPSTWYVXacgtuacgt-_Z*";
        for (int i = 0; order[i]; ++i) t[(unsigned char)order[i]] = (uint8_t)i;
        t[(unsigned char)'B'] = t[(unsigned char)'N'];     // :544
        t[(unsigned char)'Z'] = t[(unsigned char)'Q'];     // :546
        init = true;
    }
    return t;
}

// residue_types.rs:572-580
inline uint8_t standard_letter_to_index(unsigned char letter) {
    const uint8_t x = standard_letter_table()[letter];
    if (x == kInvalidLetter) throw std::invalid_argument(std::string("InvalidOneLetterCode: ") + (char)letter);
    return x;
}

// kmers.rs:17-46: sorted, de-duplicated 5-bit-per-symbol k-mers
inline std::vector<uint32_t> generate_kmers(const std::string& seq, size_t k) {
    std::vector<uint32_t> kmers;
    if (k == 0 || k > 6 || seq.size() < k) return kmers;
    const uint32_t mask = (1u << (5 * k)) - 1u;
    uint32_t code = 0;
    kmers.reserve(seq.size() - k + 1);
    for (size_t i = 0; i < seq.size(); ++i) {
        const uint8_t x = standard_letter_to_index((unsigned char)seq[i]);
        if (x > 31) throw std::logic_error("symbol value exceeds maximum allowed value 31");
        code = ((code << 5) | x) & mask;
        if (i + 1 >= k) kmers.push_back(code);
    }
    std::sort(kmers.begin(), kmers.end());
    kmers.erase(std::unique(kmers.begin(), kmers.end()), kmers.end());
    return kmers;
}

// kmers.rs:55-78
inline size_t count_intersection_sorted(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
    size_t i = 0, j = 0, count = 0;
    while (i < a.size() && j < b.size()) {
        if (a[i] < b[j]) ++i;
        else if (a[i] > b[j]) ++j;
        else { ++count; ++i; ++j; }
    }
    return count;
}

// kmers.rs:88-100; the subtractions are on usize (wrapping in a release build), the results f32
inline std::pair<float, float> kmer_identity_bounds(size_t different_kmers, size_t kmer_len, size_t min_seq_len) {
    if (min_seq_len == 0) return {0.0f, 0.0f};
    const size_t min_mutations = different_kmers / kmer_len + 1;
    const float upper = (float)(min_seq_len - min_mutations) / (float)min_seq_len;
    const size_t max_mutations = different_kmers + kmer_len - 1;
    const float lower = (float)(min_seq_len - max_mutations) / (float)min_seq_len;
    return {std::max(lower, 0.0f), std::min(upper, 1.0f)};
}

// kmers.rs:110-121
inline size_t suggest_word_length(float identity_level) {
    if (identity_level >= 0.95f) return 6;
    if (identity_level >= 0.90f) return 5;
    if (identity_level >= 0.85f) return 5;
    if (identity_level >= 0.80f) return 4;
    if (identity_level >= 0.75f) return 4;
    if (identity_level >= 0.70f) return 3;
    if (identity_level >= 0.60f) return 3;
    if (identity_level >= 0.50f) return 2;
    return 1;
}

// bucket_clustering.rs:77-91
struct Cluster {
    size_t representative;
    std::vector<size_t> members;
    explicit Cluster(size_t rep) : representative(rep), members(1, rep) {}
};

struct ClusteringStats { size_t above_threshold = 0, below_threshold = 0, aligned = 0; };

// bucket_clustering.rs:69-75,141-309
class BucketClustering {
  public:
    static constexpr size_t kBlock = 64;                  // inconclusive representatives per GPU batch
    static constexpr int kSetAll = 6, kSetReps = 7, kSetCand = 5;     // sets of the context this driver uses

    BucketClustering(Context& ctx, const std::vector<Sequence>& sequences, float id_level, const SubstitutionMatrix& blosum62)
        : ctx_(ctx), id_level_(id_level), word_size_(suggest_word_length(id_level)), blosum62_(blosum62) {
        const size_t n = sequences.size();
        lens_.resize(n);
        for (size_t i = 0; i < n; ++i) lens_[i] = sequences[i].len();
        sequence_order_.resize(n);
        std::iota(sequence_order_.begin(), sequence_order_.end(), (size_t)0);
        // :148-149: sort_by_key(Reverse(len)) is stable
        std::stable_sort(sequence_order_.begin(), sequence_order_.end(), [&](size_t a, size_t b) { return lens_[a] > lens_[b]; });
        kmer_sets_.reserve(n);
        for (const Sequence& s : sequences) kmer_sets_.push_back(generate_kmers(s.as_u8(), word_size_));
        ctx_.load(kSetAll, sequences);
    }

    // bucket_clustering.rs:162-169
    std::vector<Cluster> run() {
        std::vector<Cluster> singles;
        for (size_t i : sequence_order_) singles.emplace_back(i);
        return merge({}, singles);
    }

    // bucket_clustering.rs:171-204: the same chunk / pairwise-merge tree, evaluated in order
    std::vector<Cluster> run_n(size_t n_threads) {
        if (n_threads == 1) return run();
        std::vector<Cluster> singles;
        for (size_t i : sequence_order_) singles.emplace_back(i);
        if (singles.empty()) return singles;
        n_threads = std::min(std::max<size_t>(n_threads, 1), singles.size());
        const size_t chunk = (singles.size() + n_threads - 1) / n_threads;
        std::vector<std::vector<Cluster>> clusterings;
        for (size_t b = 0; b < singles.size(); b += chunk)
            clusterings.push_back(merge({}, std::vector<Cluster>(singles.begin() + b, singles.begin() + std::min(b + chunk, singles.size()))));
        while (clusterings.size() > 1) {
            std::vector<std::vector<Cluster>> next;
            for (size_t b = 0; b < clusterings.size(); b += 2) {
                if (b + 1 < clusterings.size()) next.push_back(merge(clusterings[b], clusterings[b + 1]));
                else next.push_back(clusterings[b]);
            }
            clusterings.swap(next);
        }
        return clusterings.back();
    }

    // bucket_clustering.rs:209-270
    std::vector<Cluster> merge(std::vector<Cluster> clusters1, const std::vector<Cluster>& clusters2) {
        std::vector<size_t> pending;
        std::vector<uint32_t> reps, nid;
        for (const Cluster& b : clusters2) {
            const size_t cand = b.representative;
            bool assigned = false;
            size_t pos = 0;
            while (pos < clusters1.size() && !assigned) {
                // k-mer verdicts up to the first certain hit; the inconclusive ones are collected
                pending.clear();
                long hit = -1;
                while (pos < clusters1.size() && pending.size() < kBlock) {
                    const int v = verdict(clusters1[pos].representative, cand);
                    if (v > 0) { hit = (long)pos; ++pos; break; }
                    if (v == 0) pending.push_back(pos);
                    else ++stats.below_threshold;
                    ++pos;
                }
                if (!pending.empty()) {
                    reps.clear();
                    for (size_t p : pending) reps.push_back((uint32_t)clusters1[p].representative);
                    identical(reps, (uint32_t)cand, nid);
                    stats.aligned += reps.size();
                    for (size_t k = 0; k < pending.size(); ++k) {
                        const float shorter = (float)std::min(lens_[cand], lens_[clusters1[pending[k]].representative]);
                        if ((float)nid[k] / shorter >= id_level_) {                  // :306-307, :245
                            Cluster& c = clusters1[pending[k]];
                            c.members.insert(c.members.end(), b.members.begin(), b.members.end());
                            assigned = true;
                            break;
                        }
                    }
                }
                if (!assigned && hit >= 0) {
                    ++stats.above_threshold;
                    Cluster& c = clusters1[(size_t)hit];
                    c.members.insert(c.members.end(), b.members.begin(), b.members.end());
                    assigned = true;
                }
            }
            if (!assigned) clusters1.push_back(b);
        }
        return clusters1;
    }

    ClusteringStats stats;

  private:
    // the k-mer part of sequence_identity (:272-292): +1 certainly above, -1 certainly below, 0 inconclusive
    int verdict(size_t rep, size_t cand) const {
        const size_t shared = count_intersection_sorted(kmer_sets_[cand], kmer_sets_[rep]);
        const size_t different = kmer_sets_[cand].size() > shared ? kmer_sets_[cand].size() - shared : 0;   // saturating_sub
        const auto bounds = kmer_identity_bounds(different, word_size_, std::min(lens_[cand], lens_[rep]));
        if (bounds.first >= id_level_) return 1;
        if (bounds.second < id_level_) return -1;
        return 0;
    }

    // n_identical of GlobalAligner(BLOSUM62, -11, -1) for query = each representative against template = the candidate (:296-300)
    void identical(const std::vector<uint32_t>& reps, uint32_t cand, std::vector<uint32_t>& nid) {
        ctx_.set_scoring(blosum62_, -11, -1);
        ctx_.ck(bsa_gather_sequences(ctx_.raw(), kSetAll, kSetReps, reps.data(), (uint32_t)reps.size()));
        ctx_.ck(bsa_gather_sequences(ctx_.raw(), kSetAll, kSetCand, &cand, 1));
        nid.assign(reps.size(), 0);
        uint64_t n_res = 0;
        ctx_.ck(bsa_align_all_pairs(ctx_.raw(), kSetReps, kSetCand, nullptr, 0, 1, BSA_WANT_IDENTICAL, nullptr, nid.data(), &n_res));
        if (n_res != reps.size()) throw std::logic_error("bsa_align_all_pairs returned an unexpected number of results");
    }

    Context& ctx_;
    float id_level_;
    size_t word_size_;
    const SubstitutionMatrix& blosum62_;
    std::vector<size_t> lens_, sequence_order_;
    std::vector<std::vector<uint32_t>> kmer_sets_;
};

// bucket_clustering.rs:33-41 / :60-67 -> clusters as lists of sequence indices (members in the reference's order)
inline std::vector<std::vector<size_t>> bucket_clustering_n(Context& ctx, const std::vector<Sequence>& sequences, float id_level,
                                                            size_t n_threads, const SubstitutionMatrix& blosum62) {
    BucketClustering bc(ctx, sequences, id_level, blosum62);
    std::vector<std::vector<size_t>> out;
    for (const Cluster& c : bc.run_n(n_threads)) out.push_back(c.members);
    return out;
}
inline std::vector<std::vector<size_t>> bucket_clustering(Context& ctx, const std::vector<Sequence>& sequences, float id_level,
                                                          const SubstitutionMatrix& blosum62) {
    return bucket_clustering_n(ctx, sequences, id_level, 1, blosum62);
}

}  // namespace bioshell_seq
