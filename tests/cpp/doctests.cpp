// The reference's own doctests (aligner.rs:68-83,134-149,199-215,274-289,335-350; seq.rs:96-104) written
// against the C++ mirror, reading like the Rust originals.  Exit code 0 = all assertions hold.
#include <cstdio>
#include <cstdlib>

#include "biogarden.hpp"

using biogarden::ds::Sequence;
using namespace biogarden::alignment;

#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } } while (0)

int main() {
    aligner::SequenceAligner al;
    {
        auto [score, a1, a2] = al.global_alignment(Sequence("PRTEINS"), Sequence("PRTWPSEIN"), score::blosum62, -11, -1);
        CHECK(score == 8); CHECK(a1 == Sequence("PRT---EINS")); CHECK(a2 == Sequence("PRTWPSEIN-"));
    }
    {
        auto [score, a1, a2] = al.local_alignment(Sequence("PLEASANTLY"), Sequence("MEANLY"), score::blosum62, -11, -1);
        CHECK(score == 12); CHECK(a1 == Sequence("LEAS")); CHECK(a2 == Sequence("MEAN"));
    }
    {
        auto [score, a1, a2] = al.fitting_alignment(
            Sequence("GCAAACCATAAGCCCTACGTGCCGCCTGTTTAAACTCGCGAACTGAATCTTCTGCTTCACGGTGAAAGTACCACAATGGTATCACACCCCAAGGAAAC"),
            Sequence("GCCGTCAGGCTGGTGTCCG"), score::unit, -1, -1);
        CHECK(score == 5); CHECK(a1 == Sequence("GCCCT-A--C-G-TG-CCG")); CHECK(a2 == Sequence("GCCGTCAGGCTGGTGTCCG"));
    }
    {
        auto [score, a1, a2] = al.overlap_alignment(Sequence("CTAAGGGATTCCGGTAATTAGACAG"), Sequence("ATAGACCATATGTCAGTGACTGTGTAA"), score::unit, -2, -2);
        CHECK(score == 2); CHECK(a1 == Sequence("ATTAGAC-AG")); CHECK(a2 == Sequence("AT-AGACCAT"));
    }
    {
        auto [score, a1, a2] = al.semiglobal_alignment(Sequence("TAGCACTTGGATTCTCGG"), Sequence("CAGCGTGG"), score::unit, -1, -1);
        CHECK(score == 4); CHECK(a1 == Sequence("TAGCA-CTTGGATTCTCGG")); CHECK(a2 == Sequence("---CAGCGTGG--------"));
    }
    CHECK(biogarden::analysis::seq::edit_distance(al, Sequence("ACTGGATTC"), Sequence("ACGT")) == 5);
    // seq.rs:64-73
    CHECK(biogarden::analysis::seq::hamming_distance(al, Sequence("GAGCCTACTAACGGGAT"), Sequence("CATCGTAATGACGGCCT")) == 7);
    try { biogarden::analysis::seq::hamming_distance(al, Sequence("ACGT"), Sequence("ACG")); CHECK(false); }
    catch (const biogarden::BioError& e) { CHECK(e.kind == biogarden::BioError::InvalidInputSize); }
    {   // stat.rs:138-152 and io/fasta.rs:95-136 through the native ingest
        std::vector<std::string> ids;
        auto t = biogarden::io::fasta::read_all(">a first\nACGT\nACGT\n>b\nACGTTCGA\r\n>c\nTTTTTTTT\n", &ids);
        CHECK(t.len() == 3); CHECK(ids.size() == 3 && ids[0] == "a" && ids[2] == "c");
        CHECK(t.data[0] == Sequence("ACGTACGT"));
        auto d = biogarden::analysis::stat::p_distance_matrix(al, t);
        CHECK(d.size() == 9 && d[0] == 0.0f && d[1] == 2.0f / 8.0f && d[3] == d[1] && d[2] == 6.0f / 8.0f && d[5] == 6.0f / 8.0f);
    }
    // error behaviour: Err(InvalidArgumentRange) / Err(InvalidInputSize) under the reference's conditions
    try { al.global_alignment(Sequence("AC"), Sequence("AC"), score::unit, 1, -1); CHECK(false); }
    catch (const biogarden::BioError& e) { CHECK(e.kind == biogarden::BioError::InvalidArgumentRange); }
    try { al.fitting_alignment(Sequence("AC"), Sequence("ACGT"), score::unit, -1, -1); CHECK(false); }
    catch (const biogarden::BioError& e) { CHECK(e.kind == biogarden::BioError::InvalidInputSize); }
    al.semiglobal_alignment(Sequence("ACGT"), Sequence("ACG"), score::unit, 1, 1);   // no sign check (aligner.rs:351-357)
    std::puts("cpp doctests ok");
    return 0;
}
