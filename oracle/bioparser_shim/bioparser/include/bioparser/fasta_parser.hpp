// oracle/bioparser_shim -- TEST INFRASTRUCTURE ONLY (see parser.hpp).
#ifndef BIOPARSER_SHIM_FASTA_HPP
#define BIOPARSER_SHIM_FASTA_HPP
#include "parser.hpp"

namespace bioparser {

template <class T>
class FastaParser : public Parser<T> {
public:
    explicit FastaParser(const std::string& path) : Parser<T>(path) {}
    std::vector<std::unique_ptr<T>> Parse(std::uint64_t, bool shorten_names = true) override {
        std::vector<std::unique_ptr<T>> out;
        if (this->done_) return out;
        this->done_ = true;
        std::string line, name, data;
        bool have = false;
        auto flush = [&]() {
            if (have) out.emplace_back(new T(name.c_str(), (std::uint32_t)name.size(), data.c_str(), (std::uint32_t)data.size()));
        };
        while (std::getline(this->in_, line)) {
            this->chomp(line);
            if (line.empty()) continue;
            if (line[0] == '>') {
                flush();
                name = this->short_name(line, shorten_names);
                data.clear();
                have = true;
            } else {
                if (!have) throw std::invalid_argument("[bioparser shim] not FASTA");
                data += line;
            }
        }
        flush();
        if (out.empty()) throw std::invalid_argument("[bioparser shim] empty or not FASTA");
        return out;
    }
};

}  // namespace bioparser
#endif
