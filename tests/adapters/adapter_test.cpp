// adapter_test.cpp — TEST INFRASTRUCTURE.  Drives the C++ adapters of ucoslam-cv3_b200/host/ through the REFERENCE's own
// classes: xflann::impl::LinearB200 next to xflann's Linear index behind the same IndexImpl interface, and
// uco_b200::VocabularyB200 next to fbow::Vocabulary::transform on the shipped orb.fbow.  Built by `make -C oracle ref`
// against /root/reference (xflann and fbow compile without OpenCV; cv::Mat comes from oracle/shim) into oracle/_ref/,
// run on the GPU box by tests/test_adapters_gpu.py.  Exit code 0 = every comparison was bit-identical.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <xflann/xflann.h>
#include <fbow/fbow.h>
#include "hamming_index_b200.h"
#include "bow_b200.h"
#include <sstream>
#include <map_types/keyframedatabase.h>
#include "keyframe_database_b200.h"

static int fails = 0;
#define EXPECT(c, msg) do { if (!(c)) { std::printf("FAIL %s\n", msg); fails++; } else std::printf("ok   %s\n", msg); } while (0)

int main(int argc, char** argv) {
    std::mt19937 rng(5);
    const int nt = 3000, nq = 500, nn = 10;
    std::vector<unsigned char> train(nt * 32), query(nq * 32);
    for (auto& b : train) b = rng();
    for (int q = 0; q < nq; q++) {  // queries = train rows with a few flipped bits -> real neighbours and ties exist
        std::memcpy(&query[q * 32], &train[(rng() % nt) * 32], 32);
        for (int f = 0; f < 12; f++) { unsigned bit = rng() % 256; query[q * 32 + bit / 8] ^= 1u << (bit % 8); }
    }
    xflann::Matrix T(XFLANN_8U, nt, 32, train.data()), Q(XFLANN_8U, nq, 32, query.data());
    // the reference path: Index(features, LinearParams) + search(KnnSearchParams), framematcher.cpp:213,239
    xflann::Index ref(T, xflann::LinearParams());
    xflann::impl::LinearB200 dev;
    dev.build(T, xflann::LinearParams());
    for (int sorted = 0; sorted < 2; sorted++) {
        std::vector<int> i1(nq * nn), d1(nq * nn), i2(nq * nn), d2(nq * nn);
        xflann::Matrix I1(XFLANN_32S, nq, nn, i1.data()), D1(XFLANN_32S, nq, nn, d1.data()), I2(XFLANN_32S, nq, nn, i2.data()),
            D2(XFLANN_32S, nq, nn, d2.data());
        xflann::KnnSearchParams sp(-1, sorted != 0);
        ref.search(Q, nn, I1, D1, sp);
        dev.search(Q, nn, I2, D2, sp);
        if (sorted) {  // Index::_search sorts after the implementation returns (index.cpp:91-102); replay that on the adapter's rows
            xflann::Index tmp;  // sort<> is private: use the documented behaviour (ascending distance, stable exchange sort)
            for (int r = 0; r < nq; r++)
                for (int a = 0; a < nn - 1; a++)
                    for (int b = a + 1; b < nn; b++)
                        if (d2[r * nn + b] < d2[r * nn + a]) { std::swap(d2[r * nn + a], d2[r * nn + b]); std::swap(i2[r * nn + a], i2[r * nn + b]); }
            EXPECT(d1 == d2, "LinearB200 distances == xflann Linear (sorted)");
        } else {
            EXPECT(i1 == i2 && d1 == d2, "LinearB200 rows == xflann Linear rows, heap order included");
        }
    }
    EXPECT(dev.size() == (uint32_t)nt && dev.getName() == "linear_b200", "IndexImpl metadata");

    if (argc > 1) {
        fbow::Vocabulary voc;
        voc.readFromFile(argv[1]);
        uco_b200::VocabularyB200 dvoc;
        dvoc.fromVocabulary(voc);
        const int n = 2000;
        std::vector<unsigned char> desc(n * 32);
        for (auto& b : desc) b = rng();
        cv::Mat f(n, 32, CV_8UC1, desc.data());
        for (int level : {3, 1}) {
            fbow::fBow a, b;
            fbow::fBow2 a2, b2;
            voc.transform(f, level, a, a2);
            dvoc.transform(f, level, b, b2);
            bool same = a.size() == b.size() && a2 == b2;
            for (auto ia = a.begin(), ib = b.begin(); same && ia != a.end(); ++ia, ++ib) {
                float x = ia->second, y = ib->second;
                same = ia->first == ib->first && std::memcmp(&x, &y, 4) == 0;
            }
            EXPECT(same, level == 3 ? "VocabularyB200::transform == fbow::Vocabulary::transform (level 3, orb.fbow)" : "... (level 1)");
            EXPECT(fbow::fBow::score(a, b) == fbow::fBow::score(a, a), "fBow::score unchanged");
        }
        bool threw = false;
        try { fbow::fBow a; fbow::fBow2 a2; cv::Mat e(0, 32, CV_8UC1, desc.data()); dvoc.transform(e, 3, a, a2); } catch (std::runtime_error&) { threw = true; }
        EXPECT(threw, "empty input throws like the reference");

        // the keyframe database: uco_b200::KPFrameDataBaseB200 next to the reference's own KeyFrameDataBase (keyframedatabase.cpp +
        // covisgraph.cpp compiled unchanged), same vocabulary, same frames, same covisibility graph
        {
            const int n_places = 8, views = 6, nd = 500;
            std::vector<std::vector<unsigned char>> base(n_places, std::vector<unsigned char>(nd * 32));
            for (auto& p : base) for (auto& b : p) b = rng();
            auto make_view = [&](int place, uint32_t idx) {
                ucoslam::Frame fr;
                fr.idx = idx;
                fr.desc = cv::Mat(nd, 32, CV_8UC1);
                std::memcpy(fr.desc.ptr<unsigned char>(0), base[place].data(), nd * 32);
                for (int r = 0; r < nd; r++) {
                    unsigned char* row = fr.desc.ptr<unsigned char>(r);
                    if (rng() % 4 == 0) { for (int b = 0; b < 32; b++) row[b] = rng(); continue; }
                    for (int fl = rng() % 5; fl > 0; fl--) { unsigned bit = rng() % 256; row[bit / 8] ^= 1u << (bit % 8); }
                }
                return fr;
            };
            ucoslam::KeyFrameDataBase refdb;
            refdb.loadFromFile(argv[1]);
            uco_b200::KPFrameDataBaseB200 devdb;
            devdb.loadFromFile(argv[1]);
            ucoslam::FrameSet fs_ref, fs_dev;
            ucoslam::CovisGraph covis;
            std::vector<uint32_t> ids;
            for (int v = 0; v < views; v++)
                for (int p = 0; p < n_places; p++) {
                    const uint32_t idx = 10 + 3 * (uint32_t)ids.size() + (p == 2 ? 500 : 0);
                    ucoslam::Frame fr = make_view(p, idx);
                    fs_ref[idx] = fr;
                    fs_ref[idx].bowvector = std::make_shared<fbow::fBow>();
                    fs_ref[idx].bowvector_level = std::make_shared<fbow::fBow2>();
                    fs_dev[idx] = fr;
                    fs_dev[idx].bowvector = std::make_shared<fbow::fBow>();
                    fs_dev[idx].bowvector_level = std::make_shared<fbow::fBow2>();
                    refdb.add(fs_ref[idx]);
                    devdb.add(fs_dev[idx]);
                    for (size_t k = 0; k < ids.size(); k++)
                        if ((int)(k % n_places) == p || rng() % 7 == 0) covis.createIncreaseEdge(ids[k], idx, 20 + rng() % 5);
                    ids.push_back(idx);
                }
            bool same_bow = true;
            for (uint32_t i : ids) same_bow = same_bow && *fs_ref[i].bowvector == *fs_dev[i].bowvector && *fs_ref[i].bowvector_level == *fs_dev[i].bowvector_level;
            EXPECT(same_bow, "KPFrameDataBaseB200::add computes the reference's bowvector / bowvector_level");
            EXPECT(devdb.size() == refdb.size() && devdb.isId(ids[3]) && !devdb.isId(1), "size / isId");
            auto run_queries = [&](const char* what) {
                bool same = true;
                size_t total = 0;
                for (int p = 0; p < n_places; p++)
                    for (int sorted = 0; sorted < 2; sorted++)
                        for (int e = 0; e < 2; e++) {
                            ucoslam::Frame q1 = make_view(p, 9999), q2 = q1;
                            q2.bowvector = std::make_shared<fbow::fBow>();
                            q2.bowvector_level = std::make_shared<fbow::fBow2>();
                            std::set<uint32_t> exc;
                            if (e) { exc.insert(ids[p]); exc.insert(ids[5]); exc.insert(77777); }
                            auto a = refdb.relocalizationCandidates(q1, fs_ref, covis, sorted != 0, e ? 0.5f : 0.f, exc);
                            auto b = devdb.relocalizationCandidates(q2, fs_dev, covis, sorted != 0, e ? 0.5f : 0.f, exc);
                            same = same && a == b;
                            total += a.size();
                        }
                EXPECT(same && total > 0, what);
            };
            run_queries("relocalizationCandidates == the reference's KeyFrameDataBase");
            EXPECT(devdb.score(fs_dev[ids[0]], fs_dev[ids[8]]) == refdb.score(fs_ref[ids[0]], fs_ref[ids[8]]), "score");
            for (size_t k = 1; k < ids.size(); k += 3) { refdb.del(fs_ref[ids[k]]); devdb.del(fs_dev[ids[k]]); }
            run_queries("... after deletions");
            std::stringstream s1(std::ios::in | std::ios::out | std::ios::binary), s2(std::ios::in | std::ios::out | std::ios::binary);
            refdb.toStream(s1);
            devdb.toStream_(s2);
            const std::string b1 = s1.str(), b2 = s2.str();
            EXPECT(b1.size() == b2.size() + sizeof(int) && std::memcmp(b1.data() + sizeof(int), b2.data(), b2.size()) == 0,
                   "toStream_ writes the reference's bytes (after KeyFrameDataBase's type tag)");
            uco_b200::KPFrameDataBaseB200 loaded;
            std::istringstream in(b2, std::ios::binary);
            loaded.fromStream_(in);
            EXPECT(loaded.size() == refdb.size() && loaded.getSignature() == devdb.getSignature(), "fromStream_ round trip");
            bool same = true;
            for (int p = 0; p < n_places; p++) {
                ucoslam::Frame q1 = make_view(p, 9999), q2 = q1;
                q2.bowvector = std::make_shared<fbow::fBow>();
                q2.bowvector_level = std::make_shared<fbow::fBow2>();
                same = same && refdb.relocalizationCandidates(q1, fs_ref, covis) == loaded.relocalizationCandidates(q2, fs_dev, covis);
            }
            EXPECT(same, "a database read back from the stream answers like the reference");
        }
    }
    std::printf("%s\n", fails ? "ADAPTERS FAILED" : "ADAPTERS OK");
    return fails ? 1 : 0;
}
