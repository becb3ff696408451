// text_pipeline.cpp — implementation of include/fl_compat/text.h and its C ABI (w2l_text_*,
// w2l_edit_distance).  Host code: dictionaries, target generation (lexicon spelling, surround, replabel, ASG dedup),
// Viterbi-path -> letters -> words, Levenshtein meters — the steps either side of the criterion in
// recipes/slimIPL/src/Train.cpp:236-254 (dictionary), :318-339 (target transform), :829-872 (evalOutput).
#include "fl_compat/text.h"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>

#include "w2l_b200.h"

namespace w2l {
int fail(int code, const std::string& msg);
}

namespace fl {
namespace lib {
namespace text {

namespace {
std::vector<std::string> splitWs(const std::string& line) {
  std::istringstream is(line);
  std::vector<std::string> out;
  std::string tok;
  while (is >> tok) out.push_back(tok);
  return out;
}
}  // namespace

Dictionary::Dictionary(std::istream& stream) { createFromStream(stream); }
Dictionary::Dictionary(const std::string& filename) {
  std::ifstream f(filename);
  if (!f) throw std::runtime_error("Dictionary: cannot open " + filename);
  createFromStream(f);
}
void Dictionary::createFromStream(std::istream& stream) {
  std::string line;
  while (std::getline(stream, line)) {
    if (line.empty()) continue;
    auto tkns = splitWs(line);
    if (tkns.empty()) continue;
    const int idx = (int)idx2entry_.size();
    for (const auto& tkn : tkns) addEntry(tkn, idx);  // all entries of a line share the index
  }
  if (!isContiguous()) throw std::runtime_error("Invalid dictionary format - not contiguous");
}
void Dictionary::addEntry(const std::string& entry, int idx) {
  if (entry2idx_.find(entry) != entry2idx_.end()) throw std::invalid_argument("Duplicate entry name in dictionary '" + entry + "'");
  entry2idx_[entry] = idx;
  if (idx2entry_.find(idx) == idx2entry_.end()) idx2entry_[idx] = entry;
}
void Dictionary::addEntry(const std::string& entry) {
  int idx = (int)idx2entry_.size();
  while (idx2entry_.find(idx) != idx2entry_.end()) ++idx;
  addEntry(entry, idx);
}
std::string Dictionary::getEntry(int idx) const {
  auto it = idx2entry_.find(idx);
  if (it == idx2entry_.end()) throw std::invalid_argument("Unknown index in dictionary '" + std::to_string(idx) + "'");
  return it->second;
}
int Dictionary::getIndex(const std::string& entry) const {
  auto it = entry2idx_.find(entry);
  if (it == entry2idx_.end()) {
    if (defaultIndex_ < 0) throw std::invalid_argument("Unknown entry in dictionary: '" + entry + "'");
    return defaultIndex_;
  }
  return it->second;
}
bool Dictionary::contains(const std::string& entry) const { return entry2idx_.find(entry) != entry2idx_.end(); }
bool Dictionary::isContiguous() const {
  for (size_t i = 0; i < indexSize(); ++i)
    if (idx2entry_.find((int)i) == idx2entry_.end()) return false;
  return true;
}
std::vector<int> Dictionary::mapEntriesToIndices(const std::vector<std::string>& entries) const {
  std::vector<int> out;
  out.reserve(entries.size());
  for (const auto& e : entries) out.push_back(getIndex(e));
  return out;
}
std::vector<std::string> Dictionary::mapIndicesToEntries(const std::vector<int>& indices) const {
  std::vector<std::string> out;
  out.reserve(indices.size());
  for (int i : indices) out.push_back(getEntry(i));
  return out;
}

LexiconMap loadWords(std::istream& stream, int maxWords) {
  LexiconMap lexicon;
  std::string line;
  while ((maxWords < 0 || (int)lexicon.size() < maxWords) && std::getline(stream, line)) {
    auto tk = splitWs(line);
    if (tk.empty()) continue;
    if (tk.size() < 2) throw std::runtime_error("[loadWords] Invalid line: " + line);
    std::vector<std::string> spelling(tk.begin() + 1, tk.end());
    auto& all = lexicon[tk[0]];
    if (std::find(all.begin(), all.end(), spelling) == all.end()) all.push_back(std::move(spelling));  // duplicates dropped
  }
  return lexicon;
}
LexiconMap loadWords(const std::string& filename, int maxWords) {
  std::ifstream f(filename);
  if (!f) throw std::runtime_error("[loadWords] Could not read file '" + filename + "'");
  return loadWords(f, maxWords);
}

std::vector<std::string> splitWrd(const std::string& word) {
  std::vector<std::string> tokens;
  const int len = (int)word.length();
  for (int i = 0; i < len;) {
    const unsigned char c = (unsigned char)word[i];
    int n = 1;  // UTF-8 sequence length from the lead byte
    if ((c & 0xE0) == 0xC0)
      n = 2;
    else if ((c & 0xF0) == 0xE0)
      n = 3;
    else if ((c & 0xF8) == 0xF0)
      n = 4;
    else if (c >= 0x80)
      throw std::runtime_error("splitWrd: invalid UTF-8 : " + word);
    if (i + n > len) throw std::runtime_error("splitWrd: invalid UTF-8 : " + word);
    tokens.push_back(word.substr((size_t)i, (size_t)n));
    i += n;
  }
  return tokens;
}

std::vector<int> packReplabels(const std::vector<int>& tokens, const Dictionary& dict, int maxReps) {
  if (tokens.empty() || maxReps <= 0) return tokens;
  std::vector<int> repIdx((size_t)maxReps + 1);
  for (int i = 1; i <= maxReps; ++i) repIdx[(size_t)i] = dict.getIndex("<" + std::to_string(i) + ">");
  std::vector<int> result;
  int prevToken = -1, numReps = 0;
  for (int token : tokens) {
    if (token == prevToken && numReps < maxReps) {
      ++numReps;
    } else {
      if (numReps > 0) {
        result.push_back(repIdx[(size_t)numReps]);
        numReps = 0;
      }
      result.push_back(token);
      prevToken = token;
    }
  }
  if (numReps > 0) result.push_back(repIdx[(size_t)numReps]);
  return result;
}
std::vector<int> unpackReplabels(const std::vector<int>& tokens, const Dictionary& dict, int maxReps) {
  if (tokens.empty() || maxReps <= 0) return tokens;
  std::unordered_map<int, int> repValue;
  for (int i = 1; i <= maxReps; ++i) repValue[dict.getIndex("<" + std::to_string(i) + ">")] = i;
  std::vector<int> result;
  int prevToken = -1;
  for (int token : tokens) {
    auto it = repValue.find(token);
    if (it == repValue.end()) {
      result.push_back(token);
      prevToken = token;
    } else if (prevToken != -1) {  // a replabel after a replabel (or at the start) is dropped
      result.insert(result.end(), (size_t)it->second, prevToken);
      prevToken = -1;
    }
  }
  return result;
}

}  // namespace text
}  // namespace lib

namespace pkg {
namespace speech {
using lib::text::Dictionary;
using lib::text::LexiconMap;

std::vector<std::string> wrd2Target(const std::string& word, const LexiconMap& lexicon, const Dictionary& dict, const std::string& wordSeparator,
                                    float targetSamplePct, bool fallback2LtrWordSepLeft, bool fallback2LtrWordSepRight, bool skipUnk) {
  (void)targetSamplePct;  // spelling sampling is a data-augmentation knob; the first spelling is the deterministic choice
  auto lit = lexicon.find(word);
  if (lit != lexicon.end() && !lit->second.empty()) return lit->second[0];
  std::vector<std::string> res;
  if (fallback2LtrWordSepLeft && !wordSeparator.empty()) res.push_back(wordSeparator);
  for (const auto& tkn : lib::text::splitWrd(word)) {
    if (dict.contains(tkn)) {
      res.push_back(tkn);
    } else if (!skipUnk) {
      throw std::invalid_argument("Unknown token '" + tkn + "' when falling back to letter target for the unknown word: " + word);
    }
  }
  if (fallback2LtrWordSepRight && !wordSeparator.empty()) res.push_back(wordSeparator);
  return res;
}
std::vector<std::string> wrd2Target(const std::vector<std::string>& words, const LexiconMap& lexicon, const Dictionary& dict,
                                    const std::string& wordSeparator, float targetSamplePct, bool fallback2LtrWordSepLeft,
                                    bool fallback2LtrWordSepRight, bool skipUnk) {
  std::vector<std::string> res;
  for (const auto& w : words) {
    auto t = wrd2Target(w, lexicon, dict, wordSeparator, targetSamplePct, fallback2LtrWordSepLeft, fallback2LtrWordSepRight, skipUnk);
    if (t.empty()) continue;
    res.insert(res.end(), t.begin(), t.end());
  }
  return res;
}

void uniq(std::vector<int>& in) { in.erase(std::unique(in.begin(), in.end()), in.end()); }
void dedup(std::vector<int>& in) { uniq(in); }
std::vector<int> validateIdx(std::vector<int> in, int badIdx) {
  in.erase(std::remove(in.begin(), in.end(), badIdx), in.end());
  return in;
}

std::vector<int> targetFeatures(const std::vector<std::string>& words, const Dictionary& tokenDict, const LexiconMap& lexicon,
                                const TargetGenerationConfig& config) {
  auto target = wrd2Target(words, lexicon, tokenDict, config.wordSeparator_, (float)config.targetSamplePct_, config.fallback2LtrWordSepLeft_,
                           config.fallback2LtrWordSepRight_, config.skipUnk_);
  std::vector<int> tgt = tokenDict.mapEntriesToIndices(target);
  if (!config.surround_.empty()) {
    const int idx = tokenDict.getIndex(config.surround_);
    tgt.push_back(idx);
    if (tgt.size() > 1) tgt.insert(tgt.begin(), idx);
  }
  if (config.replabel_ > 0) tgt = lib::text::packReplabels(tgt, tokenDict, config.replabel_);
  if (config.criterion_ == kAsgCriterion) dedup(tgt);
  if (config.eosToken_) tgt.push_back(tokenDict.getIndex(lib::text::kEosToken));
  return tgt;
}
std::vector<int> padTargets(const std::vector<std::vector<int>>& targets, int* maxLen) {
  size_t L = 0;
  for (const auto& t : targets) L = std::max(L, t.size());
  L = std::max<size_t>(L, 1);
  std::vector<int> out(targets.size() * L, kTargetPadValue);
  for (size_t b = 0; b < targets.size(); ++b) std::copy(targets[b].begin(), targets[b].end(), out.begin() + (long)(b * L));
  if (maxLen) *maxLen = (int)L;
  return out;
}
int getTargetSize(const int* target, int len) {
  int n = len;
  while (n > 0 && target[n - 1] < 0) --n;
  return n;
}

void remapLabels(std::vector<int>& labels, const Dictionary& dict, const std::string& surround, bool eosToken, int replabel) {
  if (eosToken) {
    const int eosIdx = dict.getIndex(lib::text::kEosToken);
    while (!labels.empty() && labels.back() == eosIdx) labels.pop_back();
  }
  if (replabel > 0) labels = lib::text::unpackReplabels(labels, dict, replabel);
  auto trimLabels = [&labels](int idx) {
    if (!labels.empty() && labels.back() == idx) labels.pop_back();
    if (!labels.empty() && labels.front() == idx) labels.erase(labels.begin());
  };
  if (dict.contains(kSilToken)) trimLabels(dict.getIndex(kSilToken));
  if (!surround.empty()) trimLabels(dict.getIndex(surround));
}
std::vector<std::string> tknIdx2Ltr(const std::vector<int>& labels, const Dictionary& d, bool useWordPiece, const std::string& wordSep) {
  std::vector<std::string> result;
  for (int id : labels) {
    const std::string token = d.getEntry(id);
    if (useWordPiece) {
      for (auto& c : lib::text::splitWrd(token)) result.push_back(std::move(c));
    } else {
      result.push_back(token);
    }
  }
  if (!result.empty() && !wordSep.empty()) {
    if (result.front() == wordSep) result.erase(result.begin());
    if (!result.empty() && result.back() == wordSep) result.pop_back();
  }
  return result;
}
std::vector<std::string> tknPrediction2Ltr(std::vector<int> tokens, const Dictionary& tokenDict, const std::string& criterion,
                                           const std::string& surround, bool eosToken, int replabel, bool useWordPiece,
                                           const std::string& wordSep) {
  if (criterion == kCtcCriterion || criterion == kAsgCriterion) uniq(tokens);
  if (criterion == kCtcCriterion) {
    const int blankIdx = tokenDict.getIndex(kBlankToken);
    tokens.erase(std::remove(tokens.begin(), tokens.end(), blankIdx), tokens.end());
  }
  tokens = validateIdx(tokens, -1);
  remapLabels(tokens, tokenDict, surround, eosToken, replabel);
  return tknIdx2Ltr(tokens, tokenDict, useWordPiece, wordSep);
}
std::vector<std::string> tknTarget2Ltr(std::vector<int> tokens, const Dictionary& tokenDict, const std::string& criterion,
                                       const std::string& surround, bool eosToken, int replabel, bool useWordPiece, const std::string& wordSep) {
  (void)criterion;
  if (tokens.empty()) return {};
  remapLabels(tokens, tokenDict, surround, eosToken, replabel);
  return tknIdx2Ltr(tokens, tokenDict, useWordPiece, wordSep);
}
std::vector<std::string> tkn2Wrd(const std::vector<std::string>& input, const std::string& wordSep) {
  std::vector<std::string> words;
  std::string cur;
  for (const auto& tkn : input) {
    if (tkn == wordSep) {
      if (!cur.empty()) {
        words.push_back(cur);
        cur.clear();
      }
    } else {
      cur += tkn;
    }
  }
  if (!cur.empty()) words.push_back(cur);
  return words;
}

}  // namespace speech
}  // namespace pkg

// ---- EditDistanceMeter -----------------------------------------------------------------------------------
void EditDistanceMeter::reset() { n_ = ndel_ = nins_ = nsub_ = 0; }
void EditDistanceMeter::add(int64_t n, int64_t ndel, int64_t nins, int64_t nsub) {
  n_ += n;
  ndel_ += ndel;
  nins_ += nins;
  nsub_ += nsub;
}
template <typename T>
EditDistanceMeter::ErrorState EditDistanceMeter::levensteinDistance(const std::vector<T>& in1, const std::vector<T>& in2) const {
  // in1 = hypothesis, in2 = reference; two rolling rows of error states; on ties: substitution/match, then deletion,
  // then insertion (the preference order does not change the total, only the split)
  const size_t m = in1.size(), n = in2.size();
  std::vector<ErrorState> prev(n + 1), cur(n + 1);
  for (size_t j = 0; j <= n; ++j) prev[j].ndel = (int64_t)j;  // empty hypothesis: every reference token deleted
  for (size_t i = 1; i <= m; ++i) {
    cur[0] = ErrorState();
    cur[0].nins = (int64_t)i;  // empty reference: every hypothesis token inserted
    for (size_t j = 1; j <= n; ++j) {
      ErrorState sub = prev[j - 1];
      if (!(in1[i - 1] == in2[j - 1])) sub.nsub += 1;
      ErrorState del = cur[j - 1];
      del.ndel += 1;
      ErrorState ins = prev[j];
      ins.nins += 1;
      ErrorState best = sub;
      if (del.sum() < best.sum()) best = del;
      if (ins.sum() < best.sum()) best = ins;
      cur[j] = best;
    }
    std::swap(prev, cur);
  }
  return prev[n];
}
void EditDistanceMeter::add(const std::vector<std::string>& output, const std::vector<std::string>& target) {
  const ErrorState e = levensteinDistance(output, target);
  add((int64_t)target.size(), e.ndel, e.nins, e.nsub);
}
void EditDistanceMeter::add(const std::vector<int>& output, const std::vector<int>& target) {
  const ErrorState e = levensteinDistance(output, target);
  add((int64_t)target.size(), e.ndel, e.nins, e.nsub);
}
double EditDistanceMeter::errorRate() const { return n_ > 0 ? 100.0 * (double)(ndel_ + nins_ + nsub_) / (double)n_ : 0.0; }
std::vector<double> EditDistanceMeter::value() const {
  const double n = n_ > 0 ? (double)n_ : 1.0;
  return {errorRate(), (double)n_, 100.0 * (double)nins_ / n, 100.0 * (double)ndel_ / n, 100.0 * (double)nsub_ / n};
}
std::vector<int64_t> EditDistanceMeter::valueRaw() const { return {ndel_ + nins_ + nsub_, n_, nins_, ndel_, nsub_}; }

}  // namespace fl

// ================================================================================================
// C ABI (include/w2l_b200.h): one handle = token dictionary (+ replabels, + blank for CTC) + lexicon + flags
// ================================================================================================
namespace {
struct TextPipeline {
  fl::lib::text::Dictionary dict;
  fl::lib::text::LexiconMap lexicon;
  std::string criterion, surround, wordsep;
  int replabel = 0;
  bool wordpiece = false;
};
template <typename F>
long long guardedText(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    w2l::fail(W2L_ERR_INVALID_ARGUMENT, e.what());
    return -1;
  }
}
std::vector<std::string> splitSpace(const char* s) {
  std::istringstream is(s ? s : "");
  std::vector<std::string> out;
  std::string t;
  while (is >> t) out.push_back(t);
  return out;
}
long long putJoined(const std::vector<std::string>& v, char* out, long long cap) {
  std::string s;
  for (size_t i = 0; i < v.size(); ++i) s += (i ? " " : "") + v[i];
  const long long need = (long long)s.size() + 1;
  if (out && cap >= need) std::memcpy(out, s.c_str(), (size_t)need);
  return need;
}
}  // namespace

extern "C" {
W2L_API void* w2l_text_create(const char* tokens_text, const char* lexicon_text, const char* criterion, int replabel, const char* surround,
                              int usewordpiece, const char* wordsep) {
  TextPipeline* h = nullptr;
  guardedText([&]() -> long long {
    auto t = std::make_unique<TextPipeline>();
    std::istringstream ts(tokens_text ? tokens_text : "");
    t->dict = fl::lib::text::Dictionary(ts);
    for (int r = 1; r <= replabel; ++r) t->dict.addEntry("<" + std::to_string(r) + ">");  // Train.cpp:245-247
    t->criterion = criterion ? criterion : "";
    if (t->criterion == fl::pkg::speech::kCtcCriterion) t->dict.addEntry(fl::pkg::speech::kBlankToken);  // blank last, :248-251
    if (lexicon_text && *lexicon_text) {
      std::istringstream ls(lexicon_text);
      t->lexicon = fl::lib::text::loadWords(ls);
    }
    t->surround = surround ? surround : "";
    t->wordsep = wordsep ? wordsep : "";
    t->replabel = replabel;
    t->wordpiece = usewordpiece != 0;
    h = t.release();
    return 0;
  });
  return h;
}
W2L_API void w2l_text_destroy(void* h) { delete static_cast<TextPipeline*>(h); }
W2L_API int w2l_text_num_classes(void* h) { return (int)static_cast<TextPipeline*>(h)->dict.indexSize(); }
// transcript (words separated by spaces) -> target token indices; returns the count (call with cap = 0 to size), -1 on error
W2L_API long long w2l_text_encode(void* h, const char* transcript, int32_t* out, long long cap) {
  return guardedText([&]() -> long long {
    auto* t = static_cast<TextPipeline*>(h);
    // the reference's own configuration (recipes/slimIPL/src/Train.cpp:296-305): skipUnk = true, letter fallback with the
    // word separator on the left for word pieces, on the right otherwise
    fl::pkg::speech::TargetGenerationConfig cfg(t->wordsep, 0, t->criterion, t->surround, false, t->replabel, true, t->wordpiece, !t->wordpiece);
    const std::vector<int> tgt = fl::pkg::speech::targetFeatures(splitSpace(transcript), t->dict, t->lexicon, cfg);
    if (out && cap >= (long long)tgt.size()) std::copy(tgt.begin(), tgt.end(), out);
    return (long long)tgt.size();
  });
}
// Viterbi path (n frame labels) -> letters, space-joined; returns the bytes needed incl. the terminator, -1 on error
W2L_API long long w2l_text_prediction2ltr(void* h, const int32_t* path, int n, char* out, long long cap) {
  return guardedText([&]() -> long long {
    auto* t = static_cast<TextPipeline*>(h);
    std::vector<int> v(path, path + n);
    return putJoined(fl::pkg::speech::tknPrediction2Ltr(v, t->dict, t->criterion, t->surround, false, t->replabel, t->wordpiece, t->wordsep), out, cap);
  });
}
// padded target row (len entries, negative = padding) -> letters, space-joined
W2L_API long long w2l_text_target2ltr(void* h, const int32_t* target, int len, char* out, long long cap) {
  return guardedText([&]() -> long long {
    auto* t = static_cast<TextPipeline*>(h);
    const int n = fl::pkg::speech::getTargetSize(target, len);
    std::vector<int> v(target, target + n);
    return putJoined(fl::pkg::speech::tknTarget2Ltr(v, t->dict, t->criterion, t->surround, false, t->replabel, t->wordpiece, t->wordsep), out, cap);
  });
}
// letters (space-joined) -> words (space-joined), split at the word separator
W2L_API long long w2l_text_ltr2wrd(void* h, const char* letters, char* out, long long cap) {
  return guardedText([&]() -> long long {
    auto* t = static_cast<TextPipeline*>(h);
    return putJoined(fl::pkg::speech::tkn2Wrd(splitSpace(letters), t->wordsep), out, cap);
  });
}
// EditDistanceMeter::add on one (hypothesis, reference) pair of space-joined token strings: out4 += {n, ndel, nins, nsub}
W2L_API int w2l_edit_distance(const char* hyp, const char* ref, long long* out4) {
  return (int)guardedText([&]() -> long long {
    fl::EditDistanceMeter m;
    m.add(splitSpace(hyp), splitSpace(ref));
    const auto r = m.valueRaw();  // {errors, n, nins, ndel, nsub}
    out4[0] += r[1];
    out4[1] += r[3];
    out4[2] += r[2];
    out4[3] += r[4];
    return 0;
  });
}
}
