// b200ctc -- host side of the n-gram model: ARPA text -> flat, relocatable tables.
//
// B200-native replacement for what the reference gets from the third-party kenlm package
// (kenlm.Model(path): reference decoder.py:1074, language_model.py:451) and from
// pygtrie.CharTrie over the unigram list (language_model.py:263).  The result is ONE
// contiguous blob (header + unigram array + n-gram hash table + vocabulary hash table +
// unigram-prefix hash set) that is uploaded to HBM as is and can be broadcast between
// GPUs with a single NCCL call.  KenLM binary files are not readable (ARPA only).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "b2c_common.h"
#include "b2c_lm.h"

inline u64 b2c_hash_bytes(const char* s, size_t n) {
    u64 h = 0;
    for (size_t i = 0; i < n; ++i)
        h = b2c_addmod61(b2c_mulmod61(h, B2C_HASH_BASE), static_cast<u64>(static_cast<unsigned char>(s[i])) + 1);
    return h;
}
inline u64 b2c_pow_bytes(size_t n) {
    u64 p = 1;
    for (size_t i = 0; i < n; ++i) p = b2c_mulmod61(p, B2C_HASH_BASE);
    return p;
}
inline u32 b2c_utf8_len(const char* s, size_t n) {
    u32 c = 0;
    for (size_t i = 0; i < n; ++i) c += ((static_cast<unsigned char>(s[i]) & 0xC0) != 0x80) ? 1u : 0u;
    return c;
}

struct B2cLmHeader {          // first bytes of the blob
    u64 magic;                // "B2CLM001"
    u64 total_bytes;
    int order;
    u32 bos_id, eos_id, n_vocab;
    int have_unigrams, n_unigrams;
    u64 off_uni, off_ngrams, ngram_mask, off_vocab, vocab_mask, off_prefix, prefix_mask;
    u64 n_ngrams_total;
    u64 counts[B2C_MAX_ORDER + 1];
};
#define B2C_LM_MAGIC 0x3130304D4C433242ull

struct B2cLmHost {
    std::vector<unsigned char> blob;
    std::string path;
    std::string error;
    // host-side dictionaries kept for the python-facing helper API (word in model, etc.)
    std::unordered_map<std::string, u32> vocab;
    std::unordered_set<std::string> unigram_set;

    const B2cLmHeader* header() const { return reinterpret_cast<const B2cLmHeader*>(blob.data()); }

    B2cLmView view(const void* base) const {   // view over a copy of the blob at `base` (host or device)
        const B2cLmHeader* h = header();
        const unsigned char* b = static_cast<const unsigned char*>(base);
        B2cLmView v;
        v.order = h->order;
        v.bos_id = h->bos_id; v.eos_id = h->eos_id; v.n_vocab = h->n_vocab;
        v.have_unigrams = h->have_unigrams; v.n_unigrams = h->n_unigrams;
        v.uni = reinterpret_cast<const B2cUni*>(b + h->off_uni);
        v.ngrams = reinterpret_cast<const B2cNgram*>(b + h->off_ngrams); v.ngram_mask = h->ngram_mask;
        v.vocab = reinterpret_cast<const B2cVocab*>(b + h->off_vocab); v.vocab_mask = h->vocab_mask;
        v.prefixes = reinterpret_cast<const u64*>(b + h->off_prefix); v.prefix_mask = h->prefix_mask;
        return v;
    }
};

static inline u64 b2c_pow2_at_least(u64 n) {
    u64 p = 16;
    while (p < n) p <<= 1;
    return p;
}

// Splits an ARPA line in place: returns number of whitespace separated fields (pointers into buf)
static inline int b2c_split_fields(char* buf, char** fields, int max_fields) {
    int n = 0;
    char* p = buf;
    while (*p && n < max_fields) {
        while (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n') ++p;
        if (!*p) break;
        fields[n++] = p;
        while (*p && *p != ' ' && *p != '\t' && *p != '\r' && *p != '\n') ++p;
        if (*p) *p++ = 0;
    }
    return n;
}

// unigrams == nullptr / n_unigrams < 0: "no unigrams given" (reference LanguageModel(unigrams=None))
static inline bool b2c_lm_build(B2cLmHost& lm, const char* arpa_path, const char* const* unigrams, long n_unigrams) {
    lm.path = arpa_path;
    FILE* fp = std::fopen(arpa_path, "rb");
    if (!fp) { lm.error = std::string("cannot open ") + arpa_path; return false; }
    struct Gram { std::vector<u32> ids; float prob, backoff; bool has_bo; };
    std::vector<std::vector<Gram>> grams(B2C_MAX_ORDER + 2);
    lm.vocab.clear();
    lm.vocab.emplace("<unk>", 0u);
    std::vector<std::string> id2word{"<unk>"};
    int order = 0, cur = 0;
    std::vector<char> line(1 << 16);
    char* fields[B2C_MAX_ORDER + 4];
    bool ok = true;
    while (std::fgets(line.data(), static_cast<int>(line.size()), fp)) {
        char* s = line.data();
        while (*s == ' ' || *s == '\t') ++s;
        if (*s == 0 || *s == '\n' || *s == '\r') continue;
        if (*s == '\\') {
            if (std::strstr(s, "-grams:")) {
                cur = std::atoi(s + 1);
                if (cur > B2C_MAX_ORDER) { lm.error = "n-gram order above the supported maximum"; ok = false; break; }
                if (cur > order) order = cur;
            } else if (std::strncmp(s, "\\end\\", 5) == 0) {
                break;
            }
            continue;
        }
        if (cur == 0) continue;  // "ngram N=count" header lines
        int nf = b2c_split_fields(s, fields, cur + 2);
        if (nf < cur + 1) continue;
        Gram g;
        g.prob = static_cast<float>(std::strtod(fields[0], nullptr));
        g.has_bo = nf > cur + 1;
        g.backoff = g.has_bo ? static_cast<float>(std::strtod(fields[cur + 1], nullptr)) : 0.0f;
        g.ids.resize(cur);
        for (int i = 0; i < cur; ++i) {
            auto it = lm.vocab.find(fields[1 + i]);
            if (it == lm.vocab.end()) {
                it = lm.vocab.emplace(fields[1 + i], static_cast<u32>(id2word.size())).first;
                id2word.emplace_back(fields[1 + i]);
            }
            g.ids[i] = it->second;
        }
        grams[cur].push_back(std::move(g));
    }
    std::fclose(fp);
    if (!ok) return false;
    if (order == 0) { lm.error = "no n-gram sections found in ARPA file"; return false; }

    const u32 n_vocab = static_cast<u32>(id2word.size());
    // ---- unigram array ------------------------------------------------------------------
    std::vector<B2cUni> uni(n_vocab, B2cUni{0.0f, 0.0f});
    std::vector<char> uni_present(n_vocab, 0);
    for (const Gram& g : grams[1]) {
        uni[g.ids[0]] = B2cUni{g.prob, g.backoff};
        uni_present[g.ids[0]] = 1;
    }
    if (!uni_present[0]) uni[0] = B2cUni{-100.0f, 0.0f};  // KenLM default for a missing <unk>
    // ---- n-gram table (orders >= 2), keys chained from the LAST word backwards -------------
    u64 n_hi = 0;
    for (int o = 2; o <= order; ++o) n_hi += grams[o].size();
    const u64 ng_size = b2c_pow2_at_least(n_hi * 2 + 16);
    std::vector<B2cNgram> table(ng_size, B2cNgram{0, 0.0f, 0.0f});
    auto key_of = [](const std::vector<u32>& ids, size_t n) {
        u64 h = b2c_ngram_start(ids[n - 1]);
        for (size_t k = 1; k < n; ++k) h = b2c_ngram_extend(h, ids[n - 1 - k]);
        return h;
    };
    auto find_slot = [&](u64 key) -> B2cNgram* {
        u64 slot = b2c_mix64(key) & (ng_size - 1);
        while (table[slot].key != 0 && table[slot].key != key) slot = (slot + 1) & (ng_size - 1);
        return &table[slot];
    };
    for (int o = 2; o <= order; ++o) {
        for (const Gram& g : grams[o]) {
            B2cNgram* e = find_slot(key_of(g.ids, g.ids.size()));
            e->key = key_of(g.ids, g.ids.size());
            e->prob = g.prob;
            e->backoff = g.backoff;
        }
    }
    // ---- extension marks: backoff == -0.0f  <=>  "no extension" (KenLM kNoExtensionBackoff) ----
    // an n-gram has an extension when its backoff is non-zero or it is the context (first n
    // words) of a longer n-gram.
    auto is_zero = [](float f) { return f == 0.0f; };
    const float neg_zero = -0.0f;
    for (u32 w = 0; w < n_vocab; ++w) if (is_zero(uni[w].backoff)) uni[w].backoff = neg_zero;
    for (u64 s = 0; s < ng_size; ++s) if (table[s].key && is_zero(table[s].backoff)) table[s].backoff = neg_zero;
    for (int o = 2; o <= order; ++o) {
        for (const Gram& g : grams[o]) {
            if (o == 2) {
                if (b2c_f32_bits(uni[g.ids[0]].backoff) == 0x80000000u) uni[g.ids[0]].backoff = 0.0f;
            } else {
                B2cNgram* e = find_slot(key_of(g.ids, g.ids.size() - 1));
                if (e->key && b2c_f32_bits(e->backoff) == 0x80000000u) e->backoff = 0.0f;
            }
        }
    }
    // ---- vocabulary hash + unigram set ------------------------------------------------------
    lm.unigram_set.clear();
    const bool have_uni = (unigrams != nullptr && n_unigrams >= 0);
    if (have_uni) {
        for (long i = 0; i < n_unigrams; ++i) {
            auto it = lm.vocab.find(unigrams[i]);
            if (it != lm.vocab.end() && it->second != 0) lm.unigram_set.insert(unigrams[i]);  // language_model.py:94-95
        }
    }
    const u64 vsize = b2c_pow2_at_least(static_cast<u64>(n_vocab) * 2 + 16);
    std::vector<B2cVocab> vtab(vsize, B2cVocab{0, 0, 0});
    for (u32 w = 1; w < n_vocab; ++w) {   // id 0 (<unk>) is "not in the model" (kenlm __contains__)
        const std::string& s = id2word[w];
        const u64 key = b2c_hash_bytes(s.data(), s.size()) + 1;
        u64 slot = b2c_mix64(key) & (vsize - 1);
        while (vtab[slot].key != 0 && vtab[slot].key != key) slot = (slot + 1) & (vsize - 1);
        if (vtab[slot].key == key) { lm.error = "word hash collision in vocabulary: " + s; return false; }
        vtab[slot].key = key;
        vtab[slot].id = w;
        vtab[slot].flags = lm.unigram_set.count(s) ? 1u : 0u;
    }
    // ---- prefix set over the (filtered) unigrams --------------------------------------------
    std::unordered_set<u64> pref;
    for (const std::string& s : lm.unigram_set) {
        u64 h = 0;
        for (size_t i = 0; i < s.size(); ++i) {
            h = b2c_addmod61(b2c_mulmod61(h, B2C_HASH_BASE), static_cast<u64>(static_cast<unsigned char>(s[i])) + 1);
            const bool boundary = (i + 1 == s.size()) || ((static_cast<unsigned char>(s[i + 1]) & 0xC0) != 0x80);
            if (boundary) pref.insert(h + 1);
        }
    }
    const u64 psize = b2c_pow2_at_least(pref.size() * 2 + 16);
    std::vector<u64> ptab(psize, 0);
    for (u64 key : pref) {
        u64 slot = b2c_mix64(key) & (psize - 1);
        while (ptab[slot] != 0) slot = (slot + 1) & (psize - 1);
        ptab[slot] = key;
    }
    // ---- assemble the blob ------------------------------------------------------------------
    auto align = [](u64 x) { return (x + 255) & ~255ull; };
    B2cLmHeader h;
    std::memset(&h, 0, sizeof(h));
    h.magic = B2C_LM_MAGIC;
    h.order = order;
    auto bos = lm.vocab.find("<s>");
    auto eos = lm.vocab.find("</s>");
    h.bos_id = bos == lm.vocab.end() ? 0 : bos->second;
    h.eos_id = eos == lm.vocab.end() ? 0 : eos->second;
    h.n_vocab = n_vocab;
    h.have_unigrams = have_uni ? 1 : 0;
    h.n_unigrams = static_cast<int>(lm.unigram_set.size());
    h.off_uni = align(sizeof(B2cLmHeader));
    h.off_ngrams = align(h.off_uni + sizeof(B2cUni) * n_vocab);
    h.ngram_mask = ng_size - 1;
    h.off_vocab = align(h.off_ngrams + sizeof(B2cNgram) * ng_size);
    h.vocab_mask = vsize - 1;
    h.off_prefix = align(h.off_vocab + sizeof(B2cVocab) * vsize);
    h.prefix_mask = psize - 1;
    h.total_bytes = align(h.off_prefix + sizeof(u64) * psize);
    h.n_ngrams_total = n_hi + grams[1].size();
    for (int o = 1; o <= order; ++o) h.counts[o] = grams[o].size();
    lm.blob.assign(h.total_bytes, 0);
    std::memcpy(lm.blob.data(), &h, sizeof(h));
    std::memcpy(lm.blob.data() + h.off_uni, uni.data(), sizeof(B2cUni) * n_vocab);
    std::memcpy(lm.blob.data() + h.off_ngrams, table.data(), sizeof(B2cNgram) * ng_size);
    std::memcpy(lm.blob.data() + h.off_vocab, vtab.data(), sizeof(B2cVocab) * vsize);
    std::memcpy(lm.blob.data() + h.off_prefix, ptab.data(), sizeof(u64) * psize);
    return true;
}
