// `troublet` drop-in: the reference's argv grammar (/root/reference/troublet/src/params.yml:1-49, load_params
// T:473-505), the TSV on stdout (T:36-41, T:233-243) and its log lines on stderr, on top of libsoupgpu's soup_troublet.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/soupgpu.h"

static void die(const std::string& msg, int code) {
    fprintf(stderr, "%s\n", msg.c_str());
    exit(code);
}

int main(int argc, char** argv) {
    struct Opt { const char* name; char sh; bool multiple; const char* help; };
    // params.yml:5-49, in the order clap prints them (by long name)
    static const Opt OPTS[] = {{"alts", 'a', false, "alt allele counts per cell in sparse matrix format out of vartrix"},
                               {"clusters", 'c', false, "cluster file output from schism"},
                               {"debug", 'b', true, "print debug info for index of cells listed"},
                               {"doublet_prior", 'd', false, "prior on doublets. Defaults to 0.5"},
                               {"doublet_threshold", 0, false, "doublet posterior threshold, defaults to 0.90"},
                               {"refs", 'r', false, "ref allele counts per cell in sparse matrix format out of vartrix"},
                               {"singlet_threshold", 0, false, "singlet posterior threshold, defaults to 0.90"}};
    static const char* USAGE = "USAGE:\n    troublet [OPTIONS] --alts <alts> --clusters <clusters>";
    std::map<std::string, std::vector<std::string>> vals;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-h" || a == "--help") {
            printf("troublet 2.4\nHaynes Heaton <whheaton@gmail.com>\nIntergenotypic doublet detection given cluster assignments and cell allele counts\n\n"
                   "%s\n\nFLAGS:\n    -h, --help       Prints help information\n    -V, --version    Prints version information\n\nOPTIONS:\n", USAGE);
            size_t width = 0;
            for (const Opt& q : OPTS) width = std::max(width, 2 * strlen(q.name) + (q.multiple ? 8 : 5));
            for (const Opt& q : OPTS) {   // "    -a, --alts <alts>    help", help column aligned as clap does
                std::string left = std::string("--") + q.name + " <" + q.name + ">" + (q.multiple ? "..." : "");
                if (q.sh) printf("    -%c, %-*s    %s\n", q.sh, (int)width, left.c_str(), q.help);
                else printf("        %-*s    %s\n", (int)width, left.c_str(), q.help);
            }
            return 0;
        }
        if (a == "-V" || a == "--version") { printf("troublet 2.4\n"); return 0; }
        const Opt* o = nullptr;
        std::string v;
        bool has = false;
        if (a.rfind("--", 0) == 0) {
            size_t eq = a.find('=');
            std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            for (const Opt& q : OPTS) if (name == q.name) o = &q;
            if (eq != std::string::npos) { v = a.substr(eq + 1); has = true; }
        } else if (a.size() >= 2 && a[0] == '-') {
            for (const Opt& q : OPTS) if (q.sh && a[1] == q.sh) o = &q;
            if (a.size() > 2) { v = a.substr(a[2] == '=' ? 3 : 2); has = true; }
        }
        if (!o) die("error: Found argument '" + a + "' which wasn't expected, or isn't valid in this context\n\n" + USAGE + "\n\nFor more information try --help", 1);
        if (!has) {
            if (i + 1 >= argc) die(std::string("error: The argument '--") + o->name + " <" + o->name + ">' requires a value but none was supplied", 1);
            v = argv[++i];
        }
        vals[o->name].push_back(v);
        while (o->multiple && i + 1 < argc && argv[i + 1][0] != '-') vals[o->name].push_back(argv[++i]);
    }
    {   // clap lists the missing required arguments only
        std::string missing;
        for (const char* req : {"alts", "clusters"})
            if (!vals.count(req)) missing += std::string("\n    --") + req + " <" + req + ">";
        if (!missing.empty()) die("error: The following required arguments were not provided:" + missing + "\n\n" + USAGE + "\n\nFor more information try --help", 1);
    }
    if (!vals.count("refs")) die("thread 'main' panicked at 'called `Option::unwrap()` on a `None` value'", 101);   // T:476
    auto num = [&](const char* name, const char* dflt) {
        const std::string s = vals.count(name) ? vals[name].back() : dflt;
        char* e = nullptr;
        const double v = strtod(s.c_str(), &e);
        if (s.empty() || *e) die("thread 'main' panicked at 'called `Result::unwrap()` on an `Err` value: ParseFloatError'", 101);
        return v;
    };
    soup_troublet_params p{0.0, num("doublet_prior", "0.5"), num("doublet_threshold", "0.9"), num("singlet_threshold", "0.9")};

    char* barcodes = nullptr;
    double* losses = nullptr;
    uint64_t n_cells = 0;
    int32_t k = 0, n_clusters = 0;
    if (soup_clusters_load(vals["clusters"].back().c_str(), &barcodes, &losses, &n_cells, &k, &n_clusters) != SOUP_OK)
        die(std::string("thread 'main' panicked at '") + soup_last_error(nullptr) + "'", 101);
    soup_mtx* m = nullptr;
    if (soup_mtx_load(vals["alts"].back().c_str(), vals["refs"].back().c_str(), 0, 0, 0, 0, SOUP_MTX_RAW, &m) != SOUP_OK)
        die(std::string("thread 'main' panicked at '") + soup_last_error(nullptr) + "'", 101);
    if (m->n_cells > n_cells) die("thread 'main' panicked at 'index out of bounds' (more cells in the matrices than lines in the clusters file)", 101);
    soup_ctx* ctx = nullptr;
    if (soup_ctx_create(0, nullptr, &ctx) != SOUP_OK) die(std::string("troublet: ") + soup_last_error(nullptr), 1);
    const uint64_t N = m->n_cells;
    std::vector<int32_t> status(N), bs(N), d1(N), d2(N);
    std::vector<double> sp(N), dp(N), ls(N), ld(N), clp(N * (uint64_t)k);
    soup_troublet_result r{status.data(), bs.data(), d1.data(), d2.data(), sp.data(), dp.data(), ls.data(), ld.data(), clp.data(), 0, 0, 0.0};
    if (soup_troublet(ctx, N, m->n_loci, m->nnz, m->row_ptr, m->locus, m->alt, m->ref, losses, k, n_clusters, &p, &r, barcodes, stderr) != SOUP_OK)
        die(std::string("thread 'main' panicked at '") + soup_last_error(ctx) + "'", 101);
    if (soup_troublet_write(stdout, barcodes, N, k, &r) != SOUP_OK) die(std::string("troublet: ") + soup_last_error(nullptr), 1);
    soup_ctx_destroy(ctx);
    soup_mtx_free(m);
    soup_free(barcodes);
    soup_free(losses);
    return 0;
}
