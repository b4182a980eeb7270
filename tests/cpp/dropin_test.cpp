// tests/cpp/dropin_test.cpp -- a caller written against the REFERENCE's interface
// (team_alignment.hpp / team_minimizers.hpp), linked against libteam_b200.so instead of the
// reference's static libraries. Known answers are the SURVEY.md section-4 table (recorded from the
// unmodified reference). Exit code 0 = all good.
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "team_alignment.hpp"
#include "team_minimizers.hpp"

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

static void align_case(const char* q, const char* t, team::AlignmentType ty, int es, const std::string& ec, unsigned etb,
                       int m = 1, int x = -1, int g = -1) {
    std::string cg = "stale";
    unsigned tb = 12345;
    const int s = team::Align(q, (unsigned)std::strlen(q), t, (unsigned)std::strlen(t), ty, m, x, g, &cg, &tb);
    CHECK(s == es); CHECK(cg == ec); CHECK(tb == etb);
    const int s2 = team::Align(q, (unsigned)std::strlen(q), t, (unsigned)std::strlen(t), ty, m, x, g);  // defaults: nullptr
    CHECK(s2 == es);
}

int main() {
    using T = team::AlignmentType;
    const std::string one_nul("1\0", 2);
    align_case("GTACC", "GATACGTTA", T::global, -1, "1M1I3M3I1M", 0);
    align_case("GTACC", "GATACGTTA", T::local, 3, "3M", 6);
    align_case("GTACC", "GATACGTTA", T::semiGlobal, 2, "5I1M1I2M2D", 0);
    align_case("TGACGTACATGGACA", "CGTACATGGA", T::global, 5, "3D9M2D1M", 0);
    align_case("TGACGTACATGGACA", "CGTACATGGA", T::local, 10, "10M", 11);
    align_case("TGACGTACATGGACA", "CGTACATGGA", T::semiGlobal, 10, "3D10M2D", 0);
    align_case("", "", T::global, 0, one_nul, 0);
    align_case("", "", T::local, 0, one_nul, 1);
    align_case("", "ACG", T::semiGlobal, 0, "3I", 0);
    align_case("AAAA", "CCCC", T::local, 0, one_nul, 2);
    align_case("A-CG", "ACG", T::global, 3, "1M1D2M", 0);
    align_case("ACGTACGTAA", "ACGTCGTTAA", T::local, 14, "4M1D2M1I3M", 11, 2, -3, -2);
    bool threw = false;
    try { team::Align("A", 1, "A", 1, static_cast<T>(7), 1, -1, -1); } catch (const std::invalid_argument& e) {
        threw = std::string(e.what()) == "Unknown AlignmentType provided.";
    }
    CHECK(threw);

    // batched form
    std::vector<team::AlignJob> jobs{{"GTACC", 5, "GATACGTTA", 9}, {"AA", 2, "A", 1}, {"", 0, "", 0}};
    std::vector<std::string> cgs; std::vector<unsigned> tbs;
    auto sc = team::AlignBatch(jobs, T::semiGlobal, 1, -1, -1, &cgs, &tbs);
    CHECK(sc.size() == 3 && sc[0] == 2 && sc[1] == 1 && sc[2] == 0);
    CHECK(cgs[0] == "5I1M1I2M2D" && cgs[1] == "1M1D" && cgs[2] == one_nul);

    // minimizers
    team::KMER fwd(true), rev(false);
    static_assert(sizeof(team::KMER) == 1, "layout must match the reference class");
    auto v = fwd.Minimize("TGACGTACATGGACA", 15, 3, 3);
    const unsigned eh[15] = {45, 45, 19, 14, 14, 14, 17, 6, 6, 6, 27, 47, 17, 17, 17};
    const unsigned ep[15] = {1, 1, 3, 4, 4, 4, 7, 8, 8, 8, 9, 10, 13, 13, 13};
    CHECK(v.size() == 15);
    for (size_t i = 0; i < v.size() && i < 15; ++i)
        CHECK(std::get<0>(v[i]) == eh[i] && std::get<1>(v[i]) == ep[i] && std::get<2>(v[i]) == true);
    CHECK(fwd.GetUniqueMinimizers().size() == 8);
    CHECK(fwd.GetMinimizerFrequencies().at(14) == 3);
    auto r = rev.Minimize("ACGTACGAC", 9, 3, 3);   // overwrites the shared side state, flag = false
    CHECK(r.size() == 9 && std::get<0>(r[7]) == 52 && std::get<1>(r[7]) == 7 && std::get<2>(r[7]) == false);
    CHECK(fwd.GetUniqueMinimizers().size() == 5);   // process-global: describes the LAST Minimize
    auto z = fwd.Minimize(std::string(20, 'G').c_str(), 20, 16, 3);
    CHECK(z.size() == 7 && z[0] == std::make_tuple(0u, 0u, false));
    CHECK(fwd.Minimize("ACG", 3, 4, 2).empty());
    CHECK(fwd.MappSeqCharPointerToBit("TGA", 3) == 45);
    CHECK(fwd.MappKmerBitToString(45, 3) == "231");
    CHECK(fwd.ReverseComplement("ACGN") == "NCGT");
    auto mb = team::MinimizeBatch({{"TGACGTACATGGACA", 15, true}, {"ACGTACGAC", 9, false}}, 3, 3);
    CHECK(mb.size() == 2 && mb[0] == v && mb[1] == r);

    std::printf(failures ? "dropin_test: %d FAILURES\n" : "dropin_test: OK\n", failures);
    return failures ? 1 : 0;
}
