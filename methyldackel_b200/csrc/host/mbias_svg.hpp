// M-bias plots: `MethylDackel mbias <ref> <bam> <prefix>` writes <prefix>_{OT,OB,CTOT,CTOB}.svg for every strand that
// has calls (makeSVGs, svg.c:302-437).  Host-only drawing, outside the accelerated path, but part of the drop-in surface:
// pipelines pick these files up.  The files must be byte-identical to the reference's, so the element order, the
// printf formats and the layout arithmetic (80 px margin, 500 px plot, tick rules of svg.c:29-160) are the specification
// followed here; plotting helpers cite the function they restate.
#pragma once
#include <string>
#include <vector>
#include "mbias_report.hpp"

namespace mdhost {

struct MbiasPlot {
    static constexpr int margin = 80, side = 500;              // `buffer` and `dim` of svg.c:306
    FILE *f; StrandHist s; int maxX = 0; double minY = 1.0, maxY = 0.0;
    double px(int x) const { return margin + ((double) side) * x / ((double) maxX); }                               // remapX, svg.c:166-168
    double py(double y) const { return margin + side - ((double) side) * (y - minY) / (maxY - minY); }               // remapY, svg.c:162-164
    bool any(int read, int i) const { return s.meth(read, i) || s.unmeth(read, i); }
    uint32_t m_at(int read, int i) const { return i < MD_MBIAS_MAXLEN ? s.meth(read, i) : 0u; }                    // the reference reads index l, zero-filled (MBias.c:32-37)
    uint32_t u_at(int read, int i) const { return i < MD_MBIAS_MAXLEN ? s.unmeth(read, i) : 0u; }

    // y range: extremes of the 99.9 % intervals, padded by 0.03, snapped to multiples of 0.05 (getMaxY/getMinY, svg.c:29-81)
    void y_range() {
        double hi = 0.0, lo = 1.0;
        for (int i = 0; i < s.l; ++i) for (int rd = 1; rd <= 2; ++rd) if (s.meth(rd, i) + s.unmeth(rd, i)) {
            const double a = mbias_ci(s.unmeth(rd, i), s.meth(rd, i), 1), b = mbias_ci(s.unmeth(rd, i), s.meth(rd, i), 0);
            if (a > hi) hi = a;
            if (b < lo) lo = b;
        }
        hi += 0.03;
        const int c = (int) ceil(100 * hi);
        hi = (5 * (c / 5) - c) ? (1 + c / 5) * 0.05 : (c / 5) * 0.05;
        if (hi > 0.8) hi = 1.0;
        lo -= 0.03;
        lo = 0.01 * (5 * (((int)(100 * lo)) / 5));
        if (lo < 0.2) lo = 0.0;
        maxY = hi; minY = lo;
    }
    int first_x(int read) const { for (int i = 0; i < s.l; ++i) if (s.unmeth(read, i) + s.meth(read, i)) return i; return s.l; }   // getMinX, svg.c:83-93
    void x_range() {                                                                                                // getMaxX, svg.c:95-109
        int i = s.l;
        for (; i > 0; --i) if (s.unmeth(1, i - 1) + s.meth(1, i - 1) || s.unmeth(2, i - 1) + s.meth(2, i - 1)) break;
        if (i % 5) i += 5 - (i % 5);
        maxX = i;
    }
    // shaded interval band: lower bounds left to right, upper bounds back (plotCI, svg.c:170-199)
    void band(int read, int from, const char *colour) const {
        fprintf(f, "<path d=\"M %f %f\n", px(from + 1), py(mbias_ci(u_at(read, from), m_at(read, from), 0)));
        for (int i = from + 1; i <= s.l; ++i) if (m_at(read, i) || u_at(read, i)) fprintf(f, "  L %f %f\n", px(i + 1), py(mbias_ci(u_at(read, i), m_at(read, i), 0)));
        for (int i = s.l - 1; i >= 0; --i) if (m_at(read, i) || u_at(read, i)) fprintf(f, "  L %f %f\n", px(i + 1), py(mbias_ci(u_at(read, i), m_at(read, i), 1)));
        fprintf(f, "Z\" fill=\"%s\" fill-opacity=\"0.2\"/>\n", colour);
    }
    // the methylation fraction itself (plotVals, svg.c:201-228)
    void curve(int read, int from, const char *colour) const {
        auto frac = [&](int i) { return m_at(read, i) / ((double)(m_at(read, i) + u_at(read, i))); };
        fprintf(f, "<path d=\"M %f %f\n", px(from + 1), py(frac(from)));
        for (int i = from + 1; i <= s.l; ++i) if (m_at(read, i) || u_at(read, i)) fprintf(f, "  L %f %f\n", px(i + 1), py(frac(i)));
        fprintf(f, "\" stroke=\"%s\" stroke-width=\"2\" fill-opacity=\"0\"/>\n", colour);
    }
};

// Writes the plots and prints the suggestion line (the reference prints it from inside makeSVGs, svg.c:423-435).
// `which`: bit 0 CpG, bit 1 CHG, bit 2 CHH (the y-axis label).  Returns false when a file could not be opened.
inline bool mbias_write_svgs(const char *opref, const uint32_t *hist, const int32_t lens[4], int which, FILE *err) {
    static const char *titles[4] = {"Original Top", "Original Bottom", "Complementary to the Original Top", "Complementary to the Original Bottom"};
    static const char *abbrevs[4] = {"OT", "OB", "CTOT", "CTOB"};
    static const char *colour[2] = {"rgb(248,118,109)", "rgb(0,191,196)"};
    const int B = MbiasPlot::margin, D = MbiasPlot::side;
    bool printing = false, ok = true;
    for (int k = 0; k < 4; ++k) {
        if (!lens[k]) continue;
        MbiasPlot P{nullptr, StrandHist{hist, k, lens[k]}};
        P.y_range(); P.x_range();
        const int from[2] = {P.first_x(1), P.first_x(2)};
        // x ticks every 5 positions, every 10 when that would be more than 7 (getXTicks, svg.c:111-146: only the first widening is reachable)
        int span = 5, nx = P.maxX / 5;
        if (nx > 7) { span = 10; nx = P.maxX / span; }
        // y ticks every 0.05 from the lower bound (getYTicks, svg.c:148-160)
        const double yspan = P.maxY - P.minY;
        int ny = (int)(1 + ceil(yspan / 0.05));
        if (yspan < 0.05) ny = 2;
        int l1, r1, l2, r2;
        mbias_thresholds(P.s, 1, &l1, &r1);
        mbias_thresholds(P.s, 2, &l2, &r2);
        FILE *f = fopen((std::string(opref) + "_" + abbrevs[k] + ".svg").c_str(), "w");
        if (f) {
            P.f = f;
            fprintf(f, "<svg height=\"%i\" width=\"%i\"\n", D + 2 * B, D + 2 * B);
            fprintf(f, "    xmlns=\"http://www.w3.org/2000/svg\"\n    xmlns:xlink=\"http://www.w3.org/1999/xlink\"\n    xmlns:ev=\"http://www.w3.org/2001/xml-events\">\n");
            fprintf(f, "<title>%s Strand</title>\n", titles[k]);
            fprintf(f, "<rect x=\"0\" y=\"0\" width=\"%i\" height=\"%i\" fill=\"white\" />\n", D + 2 * B, D + 2 * B);
            fprintf(f, "<text x=\"%i\" y=\"%i\" text-anchor=\"middle\">%s Strand</text>\n", B + (D >> 1), 20, titles[k]);
            fprintf(f, "<line x1=\"%i\" y1=\"%i\" x2=\"%i\" y2=\"%i\" stroke=\"black\" />\n", B, B, B, B + D);
            fprintf(f, "<line x1=\"%i\" y1=\"%i\" x2=\"%i\" y2=\"%i\" stroke=\"black\" />\n", B, B + D, B + D, B + D);
            fprintf(f, "<text x=\"15\" y=\"%i\" transform=\"rotate(270 15, %i)\" text-anchor=\"middle\" dominant-baseline=\"text-before-edge\">", B + (D >> 1), B + (D >> 1));
            std::string label;
            for (int c = 0; c < 3; ++c) if (which & (1 << c)) { if (!label.empty()) label += "/"; label += c == 0 ? "CpG" : c == 1 ? "CHG" : "CHH"; }
            if (!label.empty()) label += " ";
            fprintf(f, "%sMethylation %%</text>\n", label.c_str());
            fprintf(f, "<text x=\"%i\" y=\"%i\" text-anchor=\"middle\">Position along mapped read (5'->3' of + strand)</text>\n", B + (D >> 1), B + D + 40);
            fprintf(f, "<line x1=\"%i\" y1=\"%i\" x2=\"%i\" y2=\"%i\" stroke=\"black\" />\n", B, B + D, B, B + D + 5);
            fprintf(f, "<text x=\"%i\" y=\"%i\" text-anchor=\"middle\">%i</text>\n", B, B + D + 20, 0);
            for (int j = 0; j < nx; ++j) {
                const int t = (j + 1) * span; const double x = P.px(t);
                fprintf(f, "<line x1=\"%f\" y1=\"%i\" x2=\"%f\" y2=\"%i\" stroke-dasharray=\"5 5\" stroke=\"grey\" />\n", x, B, x, B + D);
                fprintf(f, "<line x1=\"%f\" y1=\"%i\" x2=\"%f\" y2=\"%i\" stroke=\"black\" />\n", x, B + D, x, B + D + 5);
                fprintf(f, "<text x=\"%f\" y=\"%i\" text-anchor=\"middle\">%i</text>\n", x, B + D + 20, t);
            }
            for (int j = 0; j < ny; ++j) {
                const double v = 0.05 * j + P.minY, y = P.py(v);
                fprintf(f, "<line x1=\"%i\" y1=\"%f\" x2=\"%i\" y2=\"%f\" stroke=\"black\" />\n", B, y, B - 5, y);
                fprintf(f, "<text x=\"%i\" y=\"%f\" text-anchor=\"middle\" dominant-baseline=\"middle\">%4.2f</text>\n", B - 25, y, v);
            }
            const bool has[2] = {from[0] < P.s.l, from[1] < P.s.l};
            for (int rd = 0; rd < 2; ++rd) if (has[rd]) P.band(rd + 1, from[rd], colour[rd]);
            for (int rd = 0; rd < 2; ++rd) if (has[rd]) P.curve(rd + 1, from[rd], colour[rd]);
            if (l1 + l2 + r1 + r2) {
                fprintf(f, "<text x=\"%i\" y=\"%i\" text-anchor=\"end\">--%s %i,%i,%i,%i</text>\n", 2 * B + D - 10, 2 * B + D - 10, abbrevs[k], l1, r1, l2, r2);
                const int th[4] = {l1, r1, l2, r2};
                for (int t = 0; t < 4; ++t) if (th[t])
                    fprintf(f, "<line x1=\"%f\" y1=\"%i\" x2=\"%f\" y2=\"%i\" stroke-dasharray=\"5 1\" stroke=\"%s\" stroke-width=\"1\" />\n", P.px(th[t]), D + B, P.px(th[t]), B, colour[t >> 1]);
            }
            for (int rd = 0; rd < 2; ++rd) if (has[rd]) {
                fprintf(f, "<rect x=\"%i\" y=\"%i\" width=\"20\" height=\"20\" fill=\"%s\" />\n", D + B + 10, (D >> 1) + B - 20 + 20 * rd, colour[rd]);
                fprintf(f, "<text x=\"%i\" y=\"%i\" text-anchor=\"start\" dominant-baseline=\"middle\">#%i</text>\n", D + B + 35, (D >> 1) + B - 10 + 20 * rd, rd + 1);
            }
            fprintf(f, "</svg>\n");
            fclose(f);
        } else { fprintf(err, "Couldn't open %s_%s.svg for writing!\n", opref, abbrevs[k]); ok = false; }
        if (!printing) fprintf(err, "Suggested inclusion options:");
        fprintf(err, " --%s %i,%i,%i,%i", abbrevs[k], l1, r1, l2, r2);
        printing = true;
    }
    if (printing) fprintf(err, "\n");
    return ok;
}

}  // namespace mdhost
