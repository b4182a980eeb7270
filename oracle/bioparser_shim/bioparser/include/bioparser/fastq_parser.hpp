// oracle/bioparser_shim -- TEST INFRASTRUCTURE ONLY (see parser.hpp).
#ifndef BIOPARSER_SHIM_FASTQ_HPP
#define BIOPARSER_SHIM_FASTQ_HPP
#include "parser.hpp"

namespace bioparser {

template <class T>
class FastqParser : public Parser<T> {
public:
    explicit FastqParser(const std::string& path) : Parser<T>(path) {}
    // Must throw on non-FASTQ input: the mapper relies on that to fall back to FASTA (team_mapper.cpp:533-556).
    std::vector<std::unique_ptr<T>> Parse(std::uint64_t, bool shorten_names = true) override {
        std::vector<std::unique_ptr<T>> out;
        if (this->done_) return out;
        this->done_ = true;
        std::string h, s, plus, q;
        while (std::getline(this->in_, h)) {
            this->chomp(h);
            if (h.empty()) continue;
            if (h[0] != '@') throw std::invalid_argument("[bioparser shim] not FASTQ");
            if (!std::getline(this->in_, s) || !std::getline(this->in_, plus) || !std::getline(this->in_, q))
                throw std::invalid_argument("[bioparser shim] truncated FASTQ record");
            this->chomp(s); this->chomp(plus); this->chomp(q);
            if (plus.empty() || plus[0] != '+' || q.size() != s.size())
                throw std::invalid_argument("[bioparser shim] malformed FASTQ record");
            const std::string name = this->short_name(h, shorten_names);
            out.emplace_back(new T(name.c_str(), (std::uint32_t)name.size(), s.c_str(), (std::uint32_t)s.size(), q.c_str(),
                                   (std::uint32_t)q.size()));
        }
        if (out.empty()) throw std::invalid_argument("[bioparser shim] empty or not FASTQ");
        return out;
    }
};

}  // namespace bioparser
#endif
