// The checks of the reference's mrmd/data/MultiHistogram.test.cpp written against the C++ mirror of data::MultiHistogram
// (include/mrmd/data/MultiHistogram.hpp -> mrmd_b200_hist_*): every TEST of that file in the same order, the gtest
// assertions turned into a count of failed comparisons.  Prints one JSON line; exit code 1 if a comparison failed.
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "action/ThermodynamicForce.hpp"
#include "data/MultiHistogram.hpp"
#include "data/Subdomain.hpp"

using namespace mrmd;
using data::MultiHistogram;

static int failures = 0;
static std::string failed;
static void expectNear(real_t a, real_t b, const char* what)
{
    // EXPECT_FLOAT_EQ: 4 ulps of single precision
    if (!(std::abs(a - b) <= 4e-7 * std::max(std::abs(a), std::abs(b)) + 1e-30))
    {
        ++failures;
        failed += std::string(failed.empty() ? "" : ",") + what;
    }
}
static void expectEq(idx_t a, idx_t b, const char* what)
{
    if (a != b)
    {
        ++failures;
        failed += std::string(failed.empty() ? "" : ",") + what;
    }
}

int main()
{
    {  // getBin
        MultiHistogram histogram("histogram", 0_r, 10_r, 10, 2);
        expectEq(histogram.getBin(-0.5_r), -1, "getBin");
        expectEq(histogram.getBin(0.5_r), 0, "getBin");
        expectEq(histogram.getBin(5.5_r), 5, "getBin");
        expectEq(histogram.getBin(10.5_r), -1, "getBin");
        // getBinPosition, consistencyBinToPosition
        expectNear(histogram.getBinPosition(0), 0.5_r, "getBinPosition");
        expectNear(histogram.getBinPosition(5), 5.5_r, "getBinPosition");
        for (idx_t idx = 0; idx < 10; ++idx) expectEq(histogram.getBin(histogram.getBinPosition(idx)), idx, "consistencyBinToPosition");
        // createGrid, consistencyCreateGridToGetBinPosition
        const auto grid = data::createGrid(histogram);
        expectNear(grid[0], 0.5_r, "createGrid");
        expectNear(grid[5], 5.5_r, "createGrid");
        expectNear(grid[0], histogram.getBinPosition(0), "consistencyCreateGrid");
        expectNear(grid[5], histogram.getBinPosition(5), "consistencyCreateGrid");
    }
    auto peak = [](real_t a, real_t b)
    {
        MultiHistogram histogram("histogram", 0_r, 10_r, 11, 2);
        std::vector<real_t> h(22, 0_r);
        h[5 * 2 + 0] = a;
        h[5 * 2 + 1] = b;
        histogram.data.fromHost(h);
        return histogram;
    };
    {  // scale
        auto histogram = peak(10_r, 5_r);
        histogram.scale(3_r);
        const auto h = histogram.data.toHost();
        for (idx_t idx = 0; idx < 10; ++idx)
        {
            expectNear(h[idx * 2 + 0], idx == 5 ? 30_r : 0_r, "scale");
            expectNear(h[idx * 2 + 1], idx == 5 ? 15_r : 0_r, "scale");
        }
        histogram.scale(std::vector<real_t>{2_r, 4_r});  // scale(ScalarView)
        expectNear(histogram.data(5, 0), 60_r, "scalePerHistogram");
        expectNear(histogram.data(5, 1), 60_r, "scalePerHistogram");
    }
    {  // make_symmetric
        MultiHistogram histogram("histogram", 0_r, 10_r, 10, 2);
        std::vector<real_t> h(20);
        for (idx_t idx = 0; idx < 10; ++idx)
        {
            h[idx * 2 + 0] = real_c(idx);
            h[idx * 2 + 1] = 10_r - real_c(idx);
        }
        histogram.data.fromHost(h);
        histogram.makeSymmetric();
        h = histogram.data.toHost();
        for (idx_t idx = 0; idx < 10; ++idx)
        {
            expectNear(h[idx * 2 + 0], 4.5_r, "make_symmetric");
            expectNear(h[idx * 2 + 1], 5.5_r, "make_symmetric");
        }
    }
    {  // gradient
        MultiHistogram histogram("histogram", 0_r, 10_r, 10, 3);
        std::vector<real_t> h(30);
        for (idx_t idx = 0; idx < 10; ++idx)
        {
            h[idx * 3 + 0] = 1_r;
            h[idx * 3 + 1] = real_c(idx);
            h[idx * 3 + 2] = 10_r - real_c(idx);
        }
        histogram.data.fromHost(h);
        const auto grad = data::gradient(histogram, false).data.toHost();
        for (idx_t idx = 0; idx < 10; ++idx)
        {
            expectNear(grad[idx * 3 + 0], 0_r, "gradient");
            expectNear(grad[idx * 3 + 1], 1_r, "gradient");
            expectNear(grad[idx * 3 + 2], -1_r, "gradient");
        }
    }
    {  // op_plusequal, op_minusequal, op_mulequal, op_divequal
        auto a = peak(10_r, 5_r);
        a += peak(10_r, 5_r);
        expectNear(a.data(5, 0), 20_r, "op_plusequal");
        expectNear(a.data(5, 1), 10_r, "op_plusequal");
        expectNear(a.data(4, 0), 0_r, "op_plusequal");
        auto b = peak(10_r, 5_r);
        b -= peak(8_r, 1_r);
        expectNear(b.data(5, 0), 2_r, "op_minusequal");
        expectNear(b.data(5, 1), 4_r, "op_minusequal");
        auto c = peak(10_r, 5_r);
        c *= peak(8_r, 2_r);
        expectNear(c.data(5, 0), 80_r, "op_mulequal");
        expectNear(c.data(5, 1), 10_r, "op_mulequal");
        MultiHistogram d("histogram", 0_r, 10_r, 11, 2), e("histogram", 0_r, 10_r, 11, 2);
        d.data.fromHost(std::vector<real_t>(22, 3_r));
        e.data.fromHost(std::vector<real_t>(22, 3_r));
        d /= e;
        for (real_t v : d.data.toHost()) expectNear(v, 1_r, "op_divequal");
    }
    {  // smoothen_symmetric, smoothen_constant
        auto histogram = peak(10_r, 5_r);
        const auto s = data::smoothen(histogram, 1_r, 3_r).data.toHost();
        for (idx_t idx = 1; idx < 6; ++idx)
        {
            expectNear(s[(5 - idx) * 2 + 0], s[(5 + idx) * 2 + 0], "smoothen_symmetric");
            expectNear(s[(5 - idx) * 2 + 1], s[(5 + idx) * 2 + 1], "smoothen_symmetric");
        }
        MultiHistogram constant("histogram", 0_r, 10_r, 11, 1);
        constant.data.fromHost(std::vector<real_t>(11, 2_r));
        for (real_t v : data::smoothen(constant, 1_r, 3_r).data.toHost()) expectNear(v, 2_r, "smoothen_constant");
    }
    {  // replace_if_bin_position: pos < 5 as the interval (-inf, 5)
        MultiHistogram histogram("histogram", 0_r, 10_r, 11, 2);
        std::vector<real_t> h(22);
        for (idx_t idx = 0; idx < 11; ++idx)
            for (idx_t k = 0; k < 2; ++k) h[idx * 2 + k] = real_c(idx * 10 + k);
        histogram.data.fromHost(h);
        mrmd_b200_pred below{};
        below.kind = MRMD_B200_PRED_INTERVAL;
        below.axis = 0;
        below.slabMin = -1e300;
        below.slabMax = 5_r;
        data::replace_if_bin_position(histogram, below, -1_r);
        h = histogram.data.toHost();
        for (idx_t idx = 0; idx < 11; ++idx)
            for (idx_t k = 0; k < 2; ++k)
                expectNear(h[idx * 2 + k], histogram.getBinPosition(idx) < 5_r ? -1_r : real_c(idx * 10 + k), "replace_if_bin_position");
    }
    {  // cumulativeMovingAverage (LJ_IdealGas.cpp:264-292 uses it): (10 * 1 + 12) / 11 = 2
        MultiHistogram average("average", 0_r, 1_r, 4, 1), current("current", 0_r, 1_r, 4, 1);
        average.data.fromHost(std::vector<real_t>(4, 1_r));
        current.data.fromHost(std::vector<real_t>(4, 12_r));
        data::cumulativeMovingAverage(average, current, 10_r);
        for (real_t v : average.data.toHost()) expectNear(v, 2_r, "cumulativeMovingAverage");
        MultiHistogram copy("copy", average);  // MultiHistogram(label, histogram)
        expectNear(copy.data(3, 0), 2_r, "copy");
    }
    {  // ThermodynamicForce::getForce() returns a data::MultiHistogram over the subdomain (ThermodynamicForce.hpp:57)
        data::Subdomain subdomain({0_r, 0_r, 0_r}, {10_r, 10_r, 10_r}, {1_r, 1_r, 1_r});
        action::ThermodynamicForce tf({1_r, 1_r}, subdomain, 1_r, {1_r, 1_r});
        std::vector<real_t> forces(static_cast<size_t>(tf.numBins() * 2));
        for (size_t k = 0; k < forces.size(); ++k) forces[k] = real_c(k);
        tf.setForce(forces);
        auto force = tf.getForce();
        expectEq(force.numBins, 10, "getForce");
        expectEq(force.numHistograms, 2, "getForce");
        expectNear(force.min, 0_r, "getForce");
        expectNear(force.max, 10_r, "getForce");
        expectNear(force.data(3, 1), 7_r, "getForce");
        expectNear(tf.getForce(1)[3], 7_r, "getForce(typeId)");
        force.scale(2_r);  // a copy: the operator's table stays
        expectNear(tf.getForce().data(3, 1), 7_r, "getForce copy");
    }
    std::printf("{\"failures\": %d, \"failed\": \"%s\"}\n", failures, failed.c_str());
    return failures == 0 ? 0 : 1;
}
