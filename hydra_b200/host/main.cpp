// hydra_b200 -- C++ host of the B200-native per-marker Gibbs hot path.
//
// Keeps hydra's command line (reference src/options.cpp:5-333), input formats (PLINK .bed/.bim/.fam, sparse
// .si?/.ss?/.sl? sets, .phen, .group, .mS; src/data.cpp:671-823, 1443-2007) and output files (.csv .bet .cpn .acu
// .mus.<task> .eps.<task> .mrk.<task> .xbet .xcpn; src/BayesRRm.cpp:2736-2877) and drives the CUDA kernels only
// through the C ABI of include/hydra_b200.h.  One process per GPU; `--tasks T` sets the number of hydra tasks that
// the reference would have been started with (`mpirun -np T`).  Everything numeric happens on the device: without a
// CUDA device the run fails (no CPU fallback); `--dry-run` only parses the options and the input files.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

#include "../../include/hydra_b200.h"

namespace {

struct Options {  // names follow the reference's Options class (src/options.hpp:20-138)
    std::string bayesType, bedFile, phenotypeFile, failureFile, quad_points, groupIndexFile, groupMixtureFile, mcmcOutDir, mcmcOutNam, sparseDir, sparseBsn,
        markerBlocksFile, covariatesFile, priorsFile, dPriorsFile;
    double tau0 = 1.0, s02c = 1.0, v0c = 3, v0L = 3, v0t = 3;   // bayesFHMPI, src/options.hpp:91-96
    bool bedToSparse = false, dryRun = false, readFromBedFile = false, readFromSparseFiles = false, restart = false;
    uint32_t numberMarkers = 0, numberIndividuals = 0, chainLength = 10000, burnin = 5000, thin = 5, save = 10, syncRate = 1,
             shuffleMarkers = 1, tasks = 1, device = 0, blocksPerRank = 1, rank = 0, world = 1;
    bool deviceSet = false;
    uint32_t seed = 0;
    bool seedSet = false;
    double thresholdFnz = 0.06;
    std::vector<double> S{0.01, 0.001, 0.0001};  // src/options.hpp:102-110
    std::vector<std::string> ignored;            // reference options that have no effect here
    std::string mcmcOut() const { return mcmcOutDir + "/" + mcmcOutNam; }
};

[[noreturn]] void fatal(const std::string &msg) {
    printf("FATAL  : %s\n", msg.c_str());
    fflush(stdout);
    exit(1);
}

#define HB(x)                                                     \
    do {                                                          \
        if ((x) != HB_OK) fatal(std::string(hb_last_error()));    \
    } while (0)

std::vector<std::string> split(const std::string &s, const char *seps) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        size_t j = s.find_first_of(seps, i);
        if (j == std::string::npos) j = s.size();
        if (j > i) out.push_back(s.substr(i, j - i));
        i = j + 1;
    }
    return out;
}

// The two prior files (src/data.cpp:2034-2061 read_group_priors, :2069-2096 read_dirichlet_priors): groups separated by ';',
// values by ','; one row per group. --groupPriorsFile: v0G, s02G; --dPriorsFile: one Dirichlet parameter per mixture component.
std::vector<double> read_prior_matrix(const std::string &file, uint32_t rows, uint32_t cols, const char *what) {
    std::ifstream in(file);
    if (!in) throw std::runtime_error(std::string("Error: can not open the ") + what + " file [" + file + "] to read.");
    const std::string text{std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>()};
    std::vector<std::string> grp;
    for (const std::string &g : split(text, ";")) if (g.find_first_not_of(" \t\r\n") != std::string::npos) grp.push_back(g);
    if (grp.size() != rows) throw std::runtime_error(std::string(what) + " [" + file + "]: " + std::to_string(grp.size()) + " groups, the run has " + std::to_string(rows));
    std::vector<double> out;
    for (uint32_t g = 0; g < rows; g++) {
        const std::vector<std::string> v = split(grp[g], ", \t\r\n");
        if (v.size() < cols) throw std::runtime_error(std::string(what) + " [" + file + "]: group " + std::to_string(g) + " has " + std::to_string(v.size()) + " values, " + std::to_string(cols) + " needed");
        for (uint32_t k = 0; k < cols; k++) out.push_back(std::stod(v[k]));
    }
    return out;
}

Options parse(int argc, const char **argv) {
    Options o;
    auto need = [&](int &i) -> const char * {
        if (i + 1 >= argc) throw std::runtime_error(std::string("option ") + argv[i] + " needs a value");
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--mpibayes") o.bayesType = need(i);
        else if (a == "--bfile") { o.readFromBedFile = true; o.bedFile = need(i); }
        else if (a == "--pheno") o.phenotypeFile = need(i);
        else if (a == "--failure") o.failureFile = need(i);
        else if (a == "--quad_points") o.quad_points = need(i);
        else if (a == "--groupIndexFile") o.groupIndexFile = need(i);
        else if (a == "--groupMixtureFile") o.groupMixtureFile = need(i);
        else if (a == "--mcmc-out-dir") o.mcmcOutDir = need(i);
        else if (a == "--mcmc-out-name") o.mcmcOutNam = need(i);
        else if (a == "--shuf-mark") o.shuffleMarkers = (uint32_t)atoi(need(i));
        else if (a == "--marker-blocks-file") o.markerBlocksFile = need(i);
        else if (a == "--sync-rate") o.syncRate = (uint32_t)atoi(need(i));
        else if (a == "--sparse-dir") { o.readFromSparseFiles = true; o.sparseDir = need(i); }
        else if (a == "--sparse-basename") o.sparseBsn = need(i);
        else if (a == "--number-markers") o.numberMarkers = (uint32_t)atoi(need(i));
        else if (a == "--number-individuals") o.numberIndividuals = (uint32_t)atoi(need(i));
        else if (a == "--chain-length") o.chainLength = (uint32_t)atoi(need(i));
        else if (a == "--burn-in") o.burnin = (uint32_t)atoi(need(i));
        else if (a == "--seed") { o.seed = (uint32_t)atoi(need(i)); o.seedSet = true; }
        else if (a == "--thin") o.thin = (uint32_t)atoi(need(i));
        else if (a == "--save") o.save = (uint32_t)atoi(need(i));
        else if (a == "--threshold-fnz") o.thresholdFnz = atof(need(i));
        else if (a == "--bed-to-sparse") o.bedToSparse = true;
        else if (a == "--blocks-per-rank") o.blocksPerRank = (uint32_t)atoi(need(i));
        else if (a == "--S") {
            o.S.clear();
            for (auto &t : split(need(i), " ,")) o.S.push_back(std::stod(t));
        }
        // additions of this host (the reference takes the task count from mpirun)
        else if (a == "--tasks") o.tasks = (uint32_t)atoi(need(i));
        else if (a == "--device") { o.device = (uint32_t)atoi(need(i)); o.deviceSet = true; }
        else if (a == "--rank") o.rank = (uint32_t)atoi(need(i));
        else if (a == "--world") o.world = (uint32_t)atoi(need(i));
        else if (a == "--dry-run") o.dryRun = true;
        // reference options outside the accelerated path: recognised, refused with a clear message
        else if (a == "--restart") o.restart = true;
        // --sparse-sync / --bed-sync pick the reference's MPI algorithm for the epsilon synchronisation (:2080-2453); the result is
        // the same sum. Here the GPUs always exchange the changed markers themselves over NVLink, so both are accepted as
        // no-ops; --ignore-xfiles concerns the reference's restart files, which this host does not read (state file instead).
        else if (a == "--sparse-sync" || a == "--bed-sync" || a == "--ignore-xfiles") o.ignored.push_back(a);
        else if (a == "--covariates") o.covariatesFile = need(i);   // src/options.cpp:286-290
        else if (a == "--groupPriorsFile") o.priorsFile = need(i);   // src/options.cpp:278-285
        else if (a == "--dPriorsFile") o.dPriorsFile = need(i);
        else if (a == "--tau0") o.tau0 = atof(need(i));               // src/options.cpp:116-135
        else if (a == "--v0c") o.v0c = atof(need(i));
        else if (a == "--s02c") o.s02c = atof(need(i));
        else if (a == "--v0L") o.v0L = atof(need(i));
        else if (a == "--v0t") o.v0t = atof(need(i));
        else if (a == "--check-RAM")
            throw std::runtime_error("option \"" + a + "\" of hydra is not supported by hydra_b200 yet (see DESIGN.md, out of scope)");
        else
            throw std::runtime_error("\nError: invalid option \"" + a + "\".\n");  // src/options.cpp:292-296
    }
    if (!o.bedToSparse) {  // src/options.cpp:302-323
        if (o.mcmcOutDir.empty()) throw std::runtime_error("--mcmc-out-dir CL option has to be set!");
        if (o.mcmcOutNam.empty()) throw std::runtime_error("--mcmc-out-nam CL option has to be set!");
    }
    if (o.sparseBsn.empty() != o.sparseDir.empty())  // :330-331
        throw std::runtime_error("--sparse-dir and --sparse-basename must either be both set or unset");
    if (o.groupIndexFile.empty() != o.groupMixtureFile.empty())  // src/main.cpp:147-149
        throw std::runtime_error("--groupIndexFile and --groupMixtureFile must either be both set or unset");
    if (o.numberIndividuals == 0) throw std::runtime_error("--number-individuals has to be set (src/BayesRRm.cpp:3125-3143)");
    if (o.numberMarkers == 0) throw std::runtime_error("--number-markers has to be set (src/BayesRRm.cpp:3125-3143)");
    if (o.tasks == 0) throw std::runtime_error("--tasks must be >= 1");
    // one process per GPU: rank / world from the launcher's environment (torchrun, mpirun, srun) unless given
    auto env_u = [](std::initializer_list<const char *> names, uint32_t &dst) {
        for (const char *n : names)
            if (const char *v = getenv(n)) { dst = (uint32_t)atoi(v); return true; }
        return false;
    };
    if (o.world == 1) {
        env_u({"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "SLURM_NTASKS"}, o.world);
        if (o.world > 1) env_u({"RANK", "OMPI_COMM_WORLD_RANK", "SLURM_PROCID"}, o.rank);
    }
    if (!o.deviceSet && o.world > 1) {
        o.device = o.rank;
        env_u({"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"}, o.device);
    }
    if (o.world == 0 || o.rank >= o.world) throw std::runtime_error("bad --rank/--world");
    if (o.tasks % o.world != 0) throw std::runtime_error("--tasks must be a multiple of the number of GPUs (processes)");
    if (o.thin == 0) o.thin = 1;
    if (o.save % o.thin != 0) o.save = std::max(o.thin, o.save / o.thin * o.thin);  // :1058-1066 save is a multiple of thin
    return o;
}

size_t count_lines(const std::string &path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Error: can not open the file [" + path + "] to read.");
    size_t n = 0;
    std::string l;
    while (std::getline(in, l))
        if (!l.empty()) n++;
    return n;
}

// src/data.cpp:1806-1833: third token, "NA" -> individual dropped, its line number recorded
void read_phen(const std::string &path, uint32_t n_ind, std::vector<double> &y, std::vector<uint32_t> &na) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Error: can not open the phenotype file [" + path + "] to read.");
    std::string l;
    uint32_t line = 0;
    while (std::getline(in, l)) {
        auto t = split(l, " \t\r");
        if (t.empty()) continue;
        if (t.size() < 3) throw std::runtime_error("phenotype file [" + path + "]: line " + std::to_string(line + 1) + " has fewer than 3 columns");
        if (t[2] != "NA") y.push_back(atof(t[2].c_str()));
        else na.push_back(line);
        line++;
    }
    if (line != n_ind) throw std::runtime_error("phenotype file [" + path + "] has " + std::to_string(line) + " lines, --number-individuals is " + std::to_string(n_ind));
}

// src/data.cpp:1615-1672: phenotype and covariate files read together; "FID IID c1 c2 ..." per line; an NA phenotype or an NA in
// any covariate drops the individual; X is kept as read (row-major, individuals x covariates)
void read_phen_cov(const std::string &phen, const std::string &covf, uint32_t n_ind, std::vector<double> &y, std::vector<double> &X,
                   uint32_t &n_cov, std::vector<uint32_t> &na) {
    std::ifstream inp(phen), inc(covf);
    if (!inp) throw std::runtime_error("Error: can not open the phenotype file [" + phen + "] to read.");
    if (!inc) throw std::runtime_error("Error: can not open the covariates file [" + covf + "] to read.");
    std::string lp, lc;
    uint32_t line = 0;
    n_cov = 0;
    while (std::getline(inp, lp)) {
        auto tp = split(lp, " \t\r");
        if (tp.empty()) continue;
        if (!std::getline(inc, lc)) throw std::runtime_error("covariates file [" + covf + "] is shorter than the phenotype file");
        auto tc = split(lc, " \t\r");
        if (tp.size() < 3 || tc.size() < 3) throw std::runtime_error("phenotype / covariates files: malformed line " + std::to_string(line + 1));
        bool naC = false;
        for (size_t i = 2; i < tc.size(); i++) naC |= (tc[i] == "NA");
        if (tp[2] != "NA" && !naC) {
            if (n_cov == 0) n_cov = (uint32_t)tc.size() - 2;
            if (tc.size() - 2 != n_cov) throw std::runtime_error("covariates file [" + covf + "]: line " + std::to_string(line + 1) + " has a different number of columns");
            y.push_back(atof(tp[2].c_str()));
            for (size_t i = 2; i < tc.size(); i++) X.push_back(std::stod(tc[i]));
        } else {
            na.push_back(line);
        }
        line++;
    }
    if (line != n_ind) throw std::runtime_error("phenotype file [" + phen + "] has " + std::to_string(line) + " lines, --number-individuals is " + std::to_string(n_ind));
}

// src/data.cpp:1679-1752: phenotype, covariate and failure files read together; an NA phenotype, an NA in any covariate or a
// failure of "-9" drops the individual; X is kept as read (row-major, individuals x covariates)
void read_phen_fail_cov(const std::string &phen, const std::string &covf, const std::string &failf, uint32_t n_ind, std::vector<double> &y,
                        std::vector<double> &fail, std::vector<double> &X, uint32_t &n_cov, std::vector<uint32_t> &na) {
    std::ifstream inp(phen), inc(covf), inf(failf);
    if (!inp) throw std::runtime_error("Error: can not open the phenotype file [" + phen + "] to read.");
    if (!inc) throw std::runtime_error("Error: can not open the covariates file [" + covf + "] to read.");
    if (!inf) throw std::runtime_error("Error: can not open the failure file [" + failf + "] to read.");
    std::string lp, lc, lf;
    uint32_t line = 0;
    n_cov = 0;
    while (std::getline(inp, lp)) {
        auto tp = split(lp, " \t\r");
        if (tp.empty()) continue;
        if (!std::getline(inc, lc)) throw std::runtime_error("covariates file [" + covf + "] is shorter than the phenotype file");
        if (!std::getline(inf, lf)) throw std::runtime_error("failure file [" + failf + "] is shorter than the phenotype file");
        auto tc = split(lc, " \t\r"), tf = split(lf, " \t\r");
        if (tp.size() < 3 || tc.size() < 3 || tf.empty()) throw std::runtime_error("phenotype / covariates / failure files: malformed line " + std::to_string(line + 1));
        bool naC = false;
        for (size_t i = 2; i < tc.size(); i++) naC |= (tc[i] == "NA");
        if (tp[2] != "NA" && !naC && tf[0] != "-9") {
            if (n_cov == 0) n_cov = (uint32_t)tc.size() - 2;
            if (tc.size() - 2 != n_cov) throw std::runtime_error("covariates file [" + covf + "]: line " + std::to_string(line + 1) + " has a different number of columns");
            y.push_back(atof(tp[2].c_str()));
            fail.push_back(atof(tf[0].c_str()));
            for (size_t i = 2; i < tc.size(); i++) X.push_back(std::stod(tc[i]));
        } else {
            na.push_back(line);
        }
        line++;
    }
    if (line != n_ind) throw std::runtime_error("phenotype file [" + phen + "] has " + std::to_string(line) + " lines, --number-individuals is " + std::to_string(n_ind));
}

// src/data.cpp:1753-1803: phenotype and failure files read together; NA phenotype or failure "-9" drops the individual
void read_phen_fail(const std::string &phen, const std::string &failf, uint32_t n_ind, std::vector<double> &y, std::vector<double> &fail,
                    std::vector<uint32_t> &na) {
    std::ifstream inp(phen), inf(failf);
    if (!inp) throw std::runtime_error("Error: can not open the phenotype file [" + phen + "] to read.");
    if (!inf) throw std::runtime_error("Error: can not open the file [" + failf + "] to read.");
    std::string lp, lf;
    uint32_t line = 0;
    while (std::getline(inp, lp)) {
        auto tp = split(lp, " \t\r");
        if (tp.empty()) continue;
        if (!std::getline(inf, lf)) throw std::runtime_error("failure file [" + failf + "] is shorter than the phenotype file");
        auto tf = split(lf, " \t\r");
        if (tp.size() < 3 || tf.empty()) throw std::runtime_error("phenotype / failure files: malformed line " + std::to_string(line + 1));
        if (tp[2] != "NA" && tf[0] != "-9") { y.push_back(atof(tp[2].c_str())); fail.push_back(atof(tf[0].c_str())); }
        else na.push_back(line);
        line++;
    }
    if (line != n_ind) throw std::runtime_error("phenotype file [" + phen + "] has " + std::to_string(line) + " lines, --number-individuals is " + std::to_string(n_ind));
}

// src/data.cpp:1944-1960: one integer per marker
std::vector<int32_t> read_group_file(const std::string &path, uint32_t m) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Error: can not open the group file [" + path + "] to read.");
    std::vector<int32_t> g;
    std::string l;
    while (std::getline(in, l)) {
        auto t = split(l, " \t\r");
        if (t.empty()) continue;
        g.push_back(atoi(t.back().c_str()));
    }
    if (g.size() != m) throw std::runtime_error("group file [" + path + "] has " + std::to_string(g.size()) + " entries, --number-markers is " + std::to_string(m));
    return g;
}

// src/data.cpp:1975-2007: "a,b,c;a,b,c" -- one ';'-separated row per group, strictly positive, zero component implicit
std::vector<std::vector<double>> read_mS_file(const std::string &path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Error: can not open the mixture file [" + path + "] to read.");
    std::stringstream ss;
    ss << in.rdbuf();
    std::vector<std::vector<double>> rows;
    for (auto &r : split(ss.str(), ";\n\r")) {
        std::vector<double> row;
        for (auto &t : split(r, ", \t")) {
            const double v = std::stod(t);
            if (!(v > 0.0)) throw std::runtime_error("mixture file [" + path + "]: variances must be strictly positive");
            row.push_back(v);
        }
        if (!row.empty()) rows.push_back(row);
    }
    if (rows.empty()) throw std::runtime_error("mixture file [" + path + "] is empty");
    for (auto &r : rows)
        if (r.size() != rows[0].size()) throw std::runtime_error("mixture file [" + path + "]: all groups need the same number of components");
    return rows;
}

template <class T>
std::vector<T> read_binary(const std::string &path, size_t offset_elems, size_t count) {
    std::vector<T> v(count);
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Error: can not open [" + path + "]: " + strerror(errno));
    if (fseeko(f, (off_t)(offset_elems * sizeof(T)), SEEK_SET) != 0 || fread(v.data(), sizeof(T), count, f) != count) {
        fclose(f);
        throw std::runtime_error("Error: short read from [" + path + "]");
    }
    fclose(f);
    return v;
}

struct OutFile {
    FILE *f = nullptr;
    void open(const std::string &path) {
        f = fopen(path.c_str(), "wb");  // the reference deletes old files and creates new ones (:1269-1309)
        if (!f) fatal("cannot create output file " + path + ": " + strerror(errno));
    }
    void open_append(const std::string &path) {  // --restart: the file of the interrupted run is continued
        f = fopen(path.c_str(), "ab");
        if (!f) fatal("cannot open output file " + path + ": " + strerror(errno));
    }
    template <class T>
    void put(const T *p, size_t n) {
        if (fwrite(p, sizeof(T), n, f) != n) fatal("write failed");
    }
    ~OutFile() {
        if (f) fclose(f);
    }
};

// shared output files: every process writes its marker slice at its own offset (the reference uses MPI_File_write_at_all)
struct SharedFile {
    int fd = -1;
    void open_rw(const std::string &path, bool create) {
        fd = ::open(path.c_str(), create ? (O_RDWR | O_CREAT | O_TRUNC) : O_RDWR, 0644);
        if (fd < 0) fatal("cannot open output file " + path + ": " + strerror(errno));
    }
    void write_at(const void *p, size_t bytes, size_t offset) {
        const char *c = static_cast<const char *>(p);
        while (bytes) {
            const ssize_t w = ::pwrite(fd, c, bytes, (off_t)offset);
            if (w <= 0) fatal(std::string("pwrite failed: ") + strerror(errno));
            c += w; bytes -= (size_t)w; offset += (size_t)w;
        }
    }
    ~SharedFile() {
        if (fd >= 0) ::close(fd);
    }
};

// --restart (src/BayesRRm.cpp:842-928). The reference reads its text / binary outputs back (15 digits of the .csv, the Boost
// text state of .rng); here every process keeps one state file <out>.rst.<rank>, written at every --save point:
// u32 'HBRS', u32 iteration, u32 records written to .bet/.cpn/.acu/.mus so far, then the library's opaque chain state.
constexpr uint32_t kRstMagic = 0x53524248u;
void write_restart_file(const std::string &path, hb_ctx *ctx, uint32_t it, uint32_t n_saved) {
    size_t need = 0;
    HB(hb_brr_save_state(ctx, nullptr, 0, &need));
    std::vector<unsigned char> blob(need);
    HB(hb_brr_save_state(ctx, blob.data(), blob.size(), &need));
    {
        OutFile o;
        o.open(path + ".tmp");
        const uint32_t head[3] = {kRstMagic, it, n_saved};
        o.put(head, 3);
        o.put(blob.data(), blob.size());
    }
    if (rename((path + ".tmp").c_str(), path.c_str()) != 0) fatal("cannot publish " + path);
}
void read_restart_file(const std::string &path, hb_ctx *ctx, uint32_t &it, uint32_t &n_saved) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) fatal("--restart: cannot open " + path + " (it is written at every --save point of a run with the same options)");
    uint32_t head[3] = {0, 0, 0};
    if (fread(head, 4, 3, f) != 3 || head[0] != kRstMagic) fatal("--restart: " + path + " is not a hydra_b200 restart file");
    std::vector<unsigned char> blob;
    unsigned char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) blob.insert(blob.end(), buf, buf + n);
    fclose(f);
    HB(hb_brr_load_state(ctx, blob.data(), blob.size()));
    it = head[1]; n_saved = head[2];
}
// src/BayesRRm.cpp:842-928 / data.cpp read_mcmc_output_*: the state of the last save point from the OUTPUT files (reference layouts):
//   .csv     "it, G, sigmaG[G], sigmaE, h2, m0, G, K, pi[G*K]" per thinned iteration -> the last line of a save point (it > 0, it % save == 0)
//   .xbet / .xcpn   u32 Mtot; u32 it; f64 / i32 [Mtot]          .mus.<task>  {u32 it; f64 mu} per thinned iteration
//   .eps.<task>     u32 it; u32 N; f64[N]                        .mrk.<task>  u32 it; u32 len; i32[len]
template <class T>
static void read_at(const std::string &path, size_t off, T *dst, size_t n) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) fatal("--restart: cannot open " + path);
    if (fseek(f, (long)off, SEEK_SET) != 0 || fread(dst, sizeof(T), n, f) != n) { fclose(f); fatal("--restart: " + path + " is too short"); }
    fclose(f);
}
void restart_from_outputs(const std::string &out, hb_ctx *ctx, uint32_t Mtot, uint32_t N, uint32_t G, uint32_t K, uint32_t t_first, uint32_t TL,
                          uint32_t m_start, uint32_t m_local, const std::vector<int32_t> &blkL, uint32_t save, uint32_t &it_saved, uint32_t &n_saved) {
    std::ifstream in(out + ".csv");
    if (!in) fatal("--restart: neither " + out + ".rst.<rank> nor " + out + ".csv exists");
    std::string line, best;
    uint32_t n_lines = 0, best_lines = 0;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        n_lines++;
        const uint32_t it = (uint32_t)atoi(line.c_str());
        if (it > 0 && it % save == 0) { best = line; best_lines = n_lines; }
    }
    if (best.empty()) fatal("There is no point in restarting a chain from iteration 0 (not saved anyway) => restart your analysis from scratch");  // :869-874
    std::vector<double> v;
    for (auto &tok : split(best, ", ")) v.push_back(atof(tok.c_str()));
    if (v.size() != 2 + G + 5 + (size_t)G * K || (uint32_t)v[1] != G) fatal("--restart: " + out + ".csv does not belong to a run with these groups / mixtures");
    it_saved = (uint32_t)v[0]; n_saved = best_lines;
    std::vector<double> sigmaG(v.begin() + 2, v.begin() + 2 + G), pi(v.begin() + 2 + G + 5, v.end());
    const double sigmaE = v[2 + G];
    uint32_t head[2];
    std::vector<double> beta(m_local), mu(TL), eps(N);
    std::vector<int32_t> comp(m_local), perm(m_local);
    read_at(out + ".xbet", 0, head, 2);
    if (head[0] != Mtot || head[1] != it_saved) fatal("--restart: " + out + ".xbet holds iteration " + std::to_string(head[1]) + ", the .csv file's last save point is " + std::to_string(it_saved));
    read_at(out + ".xbet", 8 + (size_t)m_start * 8, beta.data(), m_local);
    read_at(out + ".xcpn", 0, head, 2);
    if (head[0] != Mtot || head[1] != it_saved) fatal("--restart: " + out + ".xcpn does not hold iteration " + std::to_string(it_saved));
    read_at(out + ".xcpn", 8 + (size_t)m_start * 4, comp.data(), m_local);
    size_t po = 0;
    for (uint32_t t = 0; t < TL; t++) {
        const std::string ts = std::to_string(t_first + t);
        struct __attribute__((packed)) { uint32_t it; double mu; } rec;
        read_at(out + ".mus." + ts, (size_t)(n_saved - 1) * 12, reinterpret_cast<unsigned char *>(&rec), 12);
        if (rec.it != it_saved) fatal("--restart: " + out + ".mus." + ts + " does not hold iteration " + std::to_string(it_saved) + " at record " + std::to_string(n_saved - 1));
        mu[t] = rec.mu;
        read_at(out + ".mrk." + ts, 0, head, 2);
        if (head[0] != it_saved || head[1] != (uint32_t)blkL[t_first + t]) fatal("--restart: " + out + ".mrk." + ts + " does not hold iteration " + std::to_string(it_saved));
        read_at(out + ".mrk." + ts, 8, perm.data() + po, (size_t)blkL[t_first + t]);
        po += (size_t)blkL[t_first + t];
    }
    read_at(out + ".eps." + std::to_string(t_first), 0, head, 2);
    if (head[0] != it_saved || head[1] != N) fatal("--restart: " + out + ".eps." + std::to_string(t_first) + " does not hold iteration " + std::to_string(it_saved));
    read_at(out + ".eps." + std::to_string(t_first), 8, eps.data(), N);
    HB(hb_brr_restore_outputs(ctx, it_saved + 1, sigmaG.data(), pi.data(), sigmaE, mu.data(), beta.data(), comp.data(), eps.data(), perm.data()));
    printf("RESTART: from files: %s.* files (no %s.rst.<rank>: state of the output files, fresh random streams)\n", out.c_str(), out.c_str());  // :849
    printf("RESTART: iteration_to_restart_from = %u\n", it_saved);
}

// keep the .csv lines of the iterations up to `it` (first field of every line)
void truncate_csv(const std::string &path, uint32_t it) {
    std::ifstream in(path);
    if (!in) fatal("--restart: cannot open " + path);
    std::string line, keep;
    while (std::getline(in, line)) {
        if (line.empty() || (uint32_t)atoi(line.c_str()) > it) break;
        keep += line + "\n";
    }
    in.close();
    std::ofstream o(path, std::ios::trunc);
    o << keep;
}

template <class T>
void dump_file(const std::string &path, uint32_t it, uint32_t n, const T *data) {  // .eps/.mrk: u32 it, u32 n, data[n]
    OutFile o;
    o.open(path);
    o.put(&it, 1);
    o.put(&n, 1);
    o.put(data, n);
}

}  // namespace

int main(int argc, const char **argv) {
    Options opt;
    try {
        opt = parse(argc, argv);
    } catch (const std::exception &e) {  // the reference prints the thrown message and returns (src/main.cpp:179-184)
        std::cerr << e.what() << std::endl;
        return 1;
    }
    try {
        const uint32_t Nraw = opt.numberIndividuals, Mtot = opt.numberMarkers;
        if (opt.readFromBedFile) {  // src/main.cpp:69-70: .fam / .bim only give the dimensions here
            const size_t nf = count_lines(opt.bedFile + ".fam"), nb = count_lines(opt.bedFile + ".bim");
            if (nf != Nraw) throw std::runtime_error(".fam file has " + std::to_string(nf) + " individuals, --number-individuals is " + std::to_string(Nraw));
            if (nb != Mtot) throw std::runtime_error(".bim file has " + std::to_string(nb) + " markers, --number-markers is " + std::to_string(Mtot));
        }
        if (!opt.readFromBedFile && !opt.readFromSparseFiles) throw std::runtime_error("either --bfile or --sparse-dir/--sparse-basename is needed");

        std::vector<double> y, fail, Xcov;
        uint32_t n_cov = 0;
        std::vector<uint32_t> na;
        const bool bayesW = (opt.bayesType == "bayesWMPI");
        if (!opt.bedToSparse) {
            if (opt.phenotypeFile.empty()) throw std::runtime_error("--pheno has to be set");
            if (bayesW) {
                if (opt.failureFile.empty()) throw std::runtime_error("--failure has to be set for bayesWMPI");
                if (opt.quad_points.empty()) throw std::runtime_error("--quad_points has to be set for bayesWMPI (3,5,7,9,11,13,15,17,25)");
                if (!opt.covariatesFile.empty()) {
                    read_phen_fail_cov(opt.phenotypeFile, opt.covariatesFile, opt.failureFile, Nraw, y, fail, Xcov, n_cov, na);
                    printf("INFO   : using covariate file: %s (numFixedEffect = %u)\n", opt.covariatesFile.c_str(), n_cov);
                } else {
                    read_phen_fail(opt.phenotypeFile, opt.failureFile, Nraw, y, fail, na);
                }
            } else if (!opt.covariatesFile.empty()) {
                read_phen_cov(opt.phenotypeFile, opt.covariatesFile, Nraw, y, Xcov, n_cov, na);
                printf("INFO   : using covariate file: %s (numFixedEffect = %u)\n", opt.covariatesFile.c_str(), n_cov);   // :1550, data.cpp:1664
            } else {
                read_phen(opt.phenotypeFile, Nraw, y, na);
            }
        }
        // groups and mixtures (src/BayesRRm.cpp:981-996)
        std::vector<int32_t> groups;
        std::vector<std::vector<double>> mS;
        if (!opt.groupIndexFile.empty()) {
            groups = read_group_file(opt.groupIndexFile, Mtot);
            mS = read_mS_file(opt.groupMixtureFile);
            for (auto g : groups)
                if (g < 0 || (size_t)g >= mS.size()) throw std::runtime_error("group index out of range of the mixture file");
        } else {
            mS.push_back(opt.S);
        }
        const uint32_t G = (uint32_t)mS.size(), K = (uint32_t)mS[0].size() + 1;
        const int repr = (opt.readFromBedFile && opt.readFromSparseFiles) ? HB_REPR_MIXED : (opt.readFromBedFile ? HB_REPR_BED : HB_REPR_SPARSE);

        printf("INFO   : hydra_b200: %s, N = %u (%zu NA phenotypes), M = %u, %u task(s), sync rate %u, %u group(s) x %u mixtures, input %s\n",
               opt.bedToSparse ? "bed-to-sparse" : opt.bayesType.c_str(), Nraw, na.size(), Mtot, opt.tasks, opt.syncRate, G, K - 1,
               repr == HB_REPR_MIXED ? "mixed" : (repr == HB_REPR_BED ? "bed" : "sparse"));
        for (auto &ig : opt.ignored)
            printf("INFO   : option %s has no effect here (the GPUs always exchange the changed markers themselves; restart uses <out>.rst.<rank>)\n", ig.c_str());
        const bool bayesFH = (opt.bayesType == "bayesFHMPI");
        if (!opt.bedToSparse && opt.bayesType != "bayesMPI" && !bayesW && !bayesFH)
            throw std::runtime_error("--mpibayes " + opt.bayesType + ": bayesMPI (BayesRRm), bayesFHMPI (BayesFH) and bayesWMPI (BayesW) are available in this build");
        if (bayesW && (!opt.priorsFile.empty() || !opt.dPriorsFile.empty()))
            throw std::runtime_error("--groupPriorsFile / --dPriorsFile are read by bayesMPI and bayesFHMPI (src/BayesRRm.cpp:2545-2554), not by bayesWMPI");
        // per-group priors (src/main.cpp:151-157): read with the other inputs, so that a malformed file stops a dry run too
        std::vector<double> group_priors, dirichlet_priors;
        if (!opt.priorsFile.empty()) {
            group_priors = read_prior_matrix(opt.priorsFile, G, 2, "--groupPriorsFile");
            printf("INFO   : group priors (v0G, s02G) of %u group(s) read from %s\n", G, opt.priorsFile.c_str());
        }
        if (!opt.dPriorsFile.empty()) {
            dirichlet_priors = read_prior_matrix(opt.dPriorsFile, G, K, "--dPriorsFile");
            printf("INFO   : Dirichlet parameters of %u group(s) x %u components read from %s\n", G, K, opt.dPriorsFile.c_str());
        }
        if (bayesFH) {
            if (!(opt.tau0 > 0 && opt.v0t > 0 && opt.v0c > 0 && opt.s02c > 0 && opt.v0L > 0))
                throw std::runtime_error("bayesFHMPI: --tau0, --v0t, --v0c, --s02c and --v0L must be positive");
            printf("INFO   : bayesFH hyper-parameters: tau0 %g v0t %g v0c %g s02c %g v0L %g\n", opt.tau0, opt.v0t, opt.v0c, opt.s02c, opt.v0L);
        }
        if (opt.dryRun) {
            printf("INFO   : dry run: options and input files parsed, nothing computed\n");
            return 0;
        }
        if (bayesW && repr == HB_REPR_MIXED) throw std::runtime_error("bayesWMPI reads bed or sparse input, not both (src/BayesW.cpp:1149-1192)");

        // ---- device context
        hb_config cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.device = (int32_t)opt.device;
        cfg.n_ind_raw = Nraw; cfg.n_na = (uint32_t)na.size(); cfg.na_inds = na.data();
        const uint32_t TL = opt.tasks / opt.world;
        cfg.m_total = Mtot; cfg.n_tasks_total = opt.tasks; cfg.task_first = opt.rank * TL; cfg.n_tasks_local = TL;
        if (opt.world > 1 && opt.bedToSparse) throw std::runtime_error("--bed-to-sparse runs on one GPU (one process)");
        cfg.sync_rate = opt.syncRate; cfg.n_groups = G; cfg.n_mix = K;
        cfg.repr_mode = opt.bedToSparse ? HB_REPR_SPARSE : repr;
        cfg.threshold_fnz = opt.thresholdFnz;
        cfg.reserved[0] = opt.shuffleMarkers ? 0u : 1u;
        cfg.model = bayesW ? 1u : 0u;
        hb_ctx *ctx = nullptr;
        HB(hb_create(&cfg, &ctx));
        uint32_t N = 0, m_start = 0, m_local = 0, lmax = 0;
        HB(hb_get_layout(ctx, &N, &m_start, &m_local, nullptr, nullptr, nullptr, &lmax));

        // ---- genotypes -> HBM (src/data.cpp:671-739 bed, :742-823 sparse; mixed mode reads the sparse files, :991-996)
        const size_t chunk = std::max<size_t>(1, ((size_t)256 << 20) / std::max<size_t>(1, (Nraw + 3) / 4));
        if (opt.readFromSparseFiles && !opt.bedToSparse) {
            const std::string b = opt.sparseDir + "/" + opt.sparseBsn;
            for (size_t ml = 0; ml < m_local; ml += chunk) {
                const size_t n = std::min(chunk, (size_t)m_local - ml), m0 = m_start + ml;
                std::vector<uint64_t> S[3], Ln[3];
                std::vector<uint32_t> I[3];
                const char *ext[3] = {"1", "2", "m"};
                for (int w = 0; w < 3; w++) {
                    S[w] = read_binary<uint64_t>(b + ".ss" + ext[w], m0, n);
                    Ln[w] = read_binary<uint64_t>(b + ".sl" + ext[w], m0, n);
                    const uint64_t lo = S[w][0], hi = S[w][n - 1] + Ln[w][n - 1];  // src/data.cpp:1101-1105
                    I[w] = read_binary<uint32_t>(b + ".si" + ext[w], lo, hi - lo);
                    for (auto &s : S[w]) s -= lo;  // :820-822
                }
                HB(hb_stage_sparse(ctx, (uint32_t)ml, (uint32_t)n, I[0].data(), S[0].data(), Ln[0].data(), I[1].data(), S[1].data(),
                                   Ln[1].data(), I[2].data(), S[2].data(), Ln[2].data()));
            }
        } else {
            const size_t nb = (Nraw + 3) / 4;
            for (size_t ml = 0; ml < m_local; ml += chunk) {
                const size_t n = std::min(chunk, (size_t)m_local - ml), m0 = m_start + ml;
                // 3 magic bytes are skipped, never validated (src/data.cpp:700)
                auto cols = read_binary<uint8_t>(opt.bedFile + ".bed", 3 + m0 * nb, n * nb);
                HB(hb_stage_bed(ctx, (uint32_t)ml, (uint32_t)n, cols.data()));
            }
        }
        HB(hb_stage_finalize(ctx));
        printf("INFO   : genotypes staged: %.3f GB in HBM\n", (double)hb_genotype_bytes(ctx) / 1e9);

        if (opt.bedToSparse) {  // src/BayesRRm.cpp:437-770: .dim .si? .ss? .sl?
            if (opt.sparseDir.empty()) throw std::runtime_error("--bed-to-sparse needs --sparse-dir and --sparse-basename");
            mkdir(opt.sparseDir.c_str(), 0755);
            const std::string b = opt.sparseDir + "/" + opt.sparseBsn;
            std::vector<uint32_t> n1(Mtot), n2(Mtot), nm(Mtot);
            HB(hb_marker_counts(ctx, n1.data(), n2.data(), nm.data()));
            OutFile fi[3], fs[3], fl[3];
            const char *ext[3] = {"1", "2", "m"};
            for (int w = 0; w < 3; w++) { fi[w].open(b + ".si" + ext[w]); fs[w].open(b + ".ss" + ext[w]); fl[w].open(b + ".sl" + ext[w]); }
            uint64_t base[3] = {0, 0, 0};
            for (size_t m0 = 0; m0 < Mtot; m0 += chunk) {
                const size_t n = std::min(chunk, (size_t)Mtot - m0);
                size_t t[3] = {0, 0, 0};
                for (size_t i = 0; i < n; i++) { t[0] += n1[m0 + i]; t[1] += n2[m0 + i]; t[2] += nm[m0 + i]; }
                std::vector<uint32_t> I[3];
                std::vector<uint64_t> S[3], Ln[3];
                for (int w = 0; w < 3; w++) { I[w].resize(std::max<size_t>(t[w], 1)); S[w].resize(n); Ln[w].resize(n); }
                HB(hb_export_sparse(ctx, (uint32_t)m0, (uint32_t)n, I[0].data(), S[0].data(), Ln[0].data(), I[1].data(), S[1].data(),
                                    Ln[1].data(), I[2].data(), S[2].data(), Ln[2].data()));
                for (int w = 0; w < 3; w++) {
                    for (auto &s : S[w]) s += base[w];  // absolute starts, in elements (:680-684)
                    fi[w].put(I[w].data(), t[w]); fs[w].put(S[w].data(), n); fl[w].put(Ln[w].data(), n);
                    base[w] += t[w];
                }
            }
            std::ofstream dim(b + ".dim");
            dim << N << " " << Mtot << "\n";
            printf("INFO   : wrote sparse files %s.{dim,si?,ss?,sl?}: %llu ones, %llu twos, %llu missing\n", b.c_str(),
                   (unsigned long long)base[0], (unsigned long long)base[1], (unsigned long long)base[2]);
            hb_destroy(ctx);
            return 0;
        }

        // ---- chain (src/BayesRRm.cpp:1565-2877, src/BayesW.cpp:1326-2090)
        std::vector<double> mSflat((size_t)G * K, 0.0);
        for (uint32_t g = 0; g < G; g++)
            for (uint32_t k = 1; k < K; k++) mSflat[g * K + k] = mS[g][k - 1];
        uint32_t seed = opt.seedSet ? opt.seed : (uint32_t)time(nullptr);  // multi-GPU: rank 0's value is shipped with the NCCL id (below)
        // one process per GPU: NCCL unique id (and rank 0's seed) through a file next to the outputs, then hb_comm_init (a barrier)
        const bool root = (opt.rank == 0);
        auto comm_setup = [&](const std::string &out) {
            if (opt.world > 1) {
                // NCCL unique id through a file next to the outputs (rank 0 writes, the others wait for it)
                const char *job = getenv("TORCHELASTIC_RUN_ID") ? getenv("TORCHELASTIC_RUN_ID") : (getenv("SLURM_JOB_ID") ? getenv("SLURM_JOB_ID") : (getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0"));
                const std::string idf = out + ".ncclid." + job;
                // the file carries the NCCL id and rank 0's seed: without --seed every process would otherwise take its own
                // time(0), and the hyper-parameter streams (drawn on every GPU instead of MPI_Bcast, :2585, 2705, 2731) would differ
                uint8_t id[HB_NCCL_ID_BYTES + 4];
                const time_t t_start = time(nullptr);
                if (root) {
                    unlink(idf.c_str());  // a file left by a crashed run must not be picked up by the other ranks
                    HB(hb_comm_get_unique_id(id));
                    memcpy(id + HB_NCCL_ID_BYTES, &seed, 4);
                    FILE *f = fopen((idf + ".tmp").c_str(), "wb");
                    if (!f || fwrite(id, 1, sizeof(id), f) != sizeof(id)) throw std::runtime_error("cannot write " + idf);
                    fclose(f);
                    if (rename((idf + ".tmp").c_str(), idf.c_str()) != 0) throw std::runtime_error("cannot publish " + idf);
                } else {
                    bool ok = false;
                    for (int tries = 0; tries < 1200 && !ok; tries++) {
                        struct stat st;
                        if (stat(idf.c_str(), &st) == 0 && st.st_size == (off_t)sizeof(id) && st.st_mtime >= t_start - 30) {  // the launcher starts the ranks together
                            FILE *f = fopen(idf.c_str(), "rb");
                            ok = f && fread(id, 1, sizeof(id), f) == sizeof(id);
                            if (f) fclose(f);
                        }
                        if (!ok) usleep(100000);
                    }
                    if (!ok) throw std::runtime_error("timed out waiting for the NCCL id file " + idf);
                }
                HB(hb_comm_init(ctx, id, (int)opt.rank, (int)opt.world));
                if (root) unlink(idf.c_str());
                if (!opt.seedSet) memcpy(&seed, id + HB_NCCL_ID_BYTES, 4);
            }
        };
        if (bayesW) {
            // Several GPUs (one process each, src/BayesW.cpp:1645, 1799-1835): every process samples its marker range, rank 0 writes
            // the .csv and the replicated residual, every process its slice of .bet / .cpn at the reference's offsets.
            struct stat sbw;
            if (stat(opt.mcmcOutDir.c_str(), &sbw) != 0 && system(("mkdir -p " + opt.mcmcOutDir).c_str()) != 0)
                throw std::runtime_error("could not create output directory --mcmc-out-dir " + opt.mcmcOutDir);
            const std::string out = opt.mcmcOut();
            const std::string rstf = out + ".rst." + std::to_string(opt.rank);
            OutFile csv, gam;
            SharedFile bet, cpn;
            if (root && !opt.restart) {
                bet.open_rw(out + ".bet", true); cpn.open_rw(out + ".cpn", true);
                bet.write_at(&Mtot, 4, 0); cpn.write_at(&Mtot, 4, 0);
            }
            comm_setup(out);
            HB(hb_bw_init(ctx, y.data(), fail.data(), groups.empty() ? nullptr : groups.data(), mSflat.data(), (uint32_t)atoi(opt.quad_points.c_str()), seed));
            if (n_cov) HB(hb_bw_set_covariates(ctx, Xcov.data(), n_cov));   // fixed effects by ARMS, src/BayesW.cpp:1366-1413
            uint32_t it_first = 0, n_saved = 0;
            if (opt.restart) {  // as for BayesRRm: state file of the last --save point, outputs cut back to it
                uint32_t it_saved = 0;
                read_restart_file(rstf, ctx, it_saved, n_saved);
                it_first = it_saved + 1;
                if (root) {
                    truncate_csv(out + ".csv", it_saved);
                    if (truncate((out + ".bet").c_str(), 4 + (off_t)n_saved * (4 + (off_t)Mtot * 8)) != 0 ||
                        truncate((out + ".cpn").c_str(), 4 + (off_t)n_saved * (4 + (off_t)Mtot * 4)) != 0)
                        fatal("--restart: cannot cut the .bet/.cpn files back to the restart point");
                    csv.open_append(out + ".csv");
                    if (n_cov) { truncate_csv(out + ".gam", it_saved); gam.open_append(out + ".gam"); }
                    printf("INFO   : restarting after iteration %u (%u records in .bet/.cpn)\n", it_saved, n_saved);
                }
                const uint64_t v[2] = {it_saved, n_saved};   // also the barrier between rank 0's truncation and the other ranks' writes
                HB(hb_comm_check_equal(ctx, v, 2, "--restart: the restart point (iteration, records) of the <out>.rst.<rank> files"));
            } else if (root) {
                csv.open(out + ".csv");
                if (n_cov) gam.open(out + ".gam");
            }
            if (!root || opt.restart) { bet.open_rw(out + ".bet", false); cpn.open_rw(out + ".cpn", false); }
            std::vector<double> gamv(n_cov);
            std::vector<int32_t> xiv(n_cov);
            std::vector<double> beta(m_local), sigmaG(G), pi((size_t)G * K), bsq(G), eps(N);
            std::vector<int32_t> comp(m_local), cass((size_t)G * K), m0(G);
            double tot_ms = 0.0;
            for (uint32_t it = it_first; it < opt.chainLength; it++) {
                hb_bw_iter_out io;
                HB(hb_bw_iteration(ctx, nullptr, &io));
                tot_ms += io.iter_ms;
                double mu = 0, alpha = 0;
                HB(hb_bw_get_hyper(ctx, sigmaG.data(), pi.data(), &mu, &alpha, bsq.data(), cass.data(), m0.data()));
                double sG = 0.0; int m0s = 0;
                for (uint32_t g = 0; g < G; g++) { sG += sigmaG[g]; m0s += m0[g]; }
                if (root) printf("%u. %d; %.7g; %.7g; %.7g\n", it, m0s, mu, alpha, sG);  // src/BayesW.cpp:1909-1911
                if (it % opt.thin == 0) {
                    char buff[65536];
                    if (root) {
                        int n = snprintf(buff, sizeof(buff), "%5d, %20.15f, %20.15f, %20.15f, %20.15f, %7d, %7d, %2d", (int)it, mu, sG, alpha,
                                         sG / (sG + 9.86960440109 / (6 * alpha * alpha)), m0s, (int)G, (int)K);  // :1942
                        for (uint32_t g = 0; g < G; g++) n += snprintf(buff + n, sizeof(buff) - n, ", %20.15f", sigmaG[g]);
                        for (size_t x = 0; x < (size_t)G * K; x++) n += snprintf(buff + n, sizeof(buff) - n, ", %20.15f", pi[x]);
                        n += snprintf(buff + n, sizeof(buff) - n, "\n");
                        csv.put(buff, (size_t)n);
                        fflush(csv.f);
                    }
                    HB(hb_brr_get_state(ctx, beta.data(), comp.data(), nullptr));
                    const size_t ob = 4 + (size_t)n_saved * (4 + (size_t)Mtot * 8), oc = 4 + (size_t)n_saved * (4 + (size_t)Mtot * 4);
                    if (root) { bet.write_at(&it, 4, ob); cpn.write_at(&it, 4, oc); }                 // :1950-1951
                    bet.write_at(beta.data(), (size_t)m_local * 8, ob + 4 + (size_t)m_start * 8);      // :1955-1957
                    cpn.write_at(comp.data(), (size_t)m_local * 4, oc + 4 + (size_t)m_start * 4);
                    if (n_cov && root) {   // "%5d, %20.17f, ..." (:1970-1980)
                        HB(hb_bw_get_gamma(ctx, gamv.data(), xiv.data()));
                        int ng = snprintf(buff, sizeof(buff), "%5d", (int)it);
                        for (uint32_t f = 0; f < n_cov; f++) ng += snprintf(buff + ng, sizeof(buff) - ng, ", %20.17f", gamv[f]);
                        ng += snprintf(buff + ng, sizeof(buff) - ng, "\n");
                        gam.put(buff, (size_t)ng);
                        fflush(gam.f);
                    }
                    n_saved++;
                }
                if (it > 0 && it % opt.save == 0) {
                    if (root) {
                        if (n_cov) { HB(hb_bw_get_gamma(ctx, gamv.data(), xiv.data())); dump_file(out + ".xiv", it, n_cov, xiv.data()); }   // :1982-1990
                        HB(hb_get_epsilon(ctx, eps.data()));
                        dump_file(out + ".eps.0", it, N, eps.data());
                    }
                    write_restart_file(rstf, ctx, it, n_saved);
                }
            }
            if (root) printf("INFO   : time to process the data: %.3f sec\n", tot_ms * 1e-3);
            hb_destroy(ctx);
            return 0;
        }
        struct stat sb;
        if (stat(opt.mcmcOutDir.c_str(), &sb) != 0 && system(("mkdir -p " + opt.mcmcOutDir).c_str()) != 0)
            throw std::runtime_error("could not create output directory --mcmc-out-dir " + opt.mcmcOutDir);
        const std::string out = opt.mcmcOut();
        // rank 0 creates the shared files (old ones are replaced, :1269-1309) before the communicator exists; hb_comm_init
        // is a barrier, after which the other processes open them
        SharedFile bet, acu, cpn, xb, xc;
        OutFile csv;
        if (root && !opt.restart) {
            csv.open(out + ".csv");
            bet.open_rw(out + ".bet", true); acu.open_rw(out + ".acu", true); cpn.open_rw(out + ".cpn", true);
            xb.open_rw(out + ".xbet", true); xc.open_rw(out + ".xcpn", true);
            bet.write_at(&Mtot, 4, 0); acu.write_at(&Mtot, 4, 0); cpn.write_at(&Mtot, 4, 0);  // :1304-1308
            xb.write_at(&Mtot, 4, 0); xc.write_at(&Mtot, 4, 0);
        }
        comm_setup(out);
        HB(hb_brr_init(ctx, y.data(), groups.empty() ? nullptr : groups.data(), mSflat.data(), nullptr, seed));  // multi-GPU: also checks that the seed is common
        if (n_cov) HB(hb_brr_set_covariates(ctx, Xcov.data(), n_cov));   // src/BayesRRm.cpp:1546-1560, 2648-2681
        if (!group_priors.empty() || !dirichlet_priors.empty())              // :2545-2554
            HB(hb_brr_set_group_priors(ctx, group_priors.empty() ? nullptr : group_priors.data(), dirichlet_priors.empty() ? nullptr : dirichlet_priors.data()));
        if (bayesFH) {                                                       // src/BayesRRm.cpp:1125-1163
            hb_fh_config fc{opt.v0L, opt.v0t, opt.v0c, opt.s02c, opt.tau0};
            HB(hb_brr_set_fh(ctx, &fc, nullptr));
            double sc[3];
            HB(hb_brr_get_fh(ctx, sc, nullptr, nullptr, nullptr));
            if (root) printf("INFO   : bayesFH: tau0 %g v0t %g v0c %g s02c %g v0L %g; initial hypTau %.6g tau %.6g\n", opt.tau0, opt.v0t, opt.v0c, opt.s02c, opt.v0L, sc[0], sc[1]);
        }
        if (!root || opt.restart) {
            bet.open_rw(out + ".bet", false); acu.open_rw(out + ".acu", false); cpn.open_rw(out + ".cpn", false);
            xb.open_rw(out + ".xbet", false); xc.open_rw(out + ".xcpn", false);
        }
        if (root) {   // list of the files of a dump (:1244-1262), one set of per-task files per reference rank = task
            std::ofstream lst(out + ".lst");
            lst << out + ".csv" << "\n" << out + ".xbet" << "\n" << out + ".xcpn" << "\n" << out + ".acu" << "\n";
            for (uint32_t i = 0; i < opt.tasks; i++)
                for (const char *e : {".rng.", ".mrk.", ".xiv.", ".eps.", ".gam.", ".mus."}) lst << out + e + std::to_string(i) << "\n";
        }
        const uint32_t t_first = opt.rank * TL;
        std::vector<OutFile> mus(TL);
        uint32_t it_first = 0, n_saved = 0;
        if (opt.restart) {
            // continue the chain after the last --save point of the interrupted run: state back into the library, output
            // files cut back to that iteration (the reference does the same from its own files, :842-928)
            uint32_t it_saved = 0;
            struct stat srst;
            if (stat((out + ".rst." + std::to_string(opt.rank)).c_str(), &srst) == 0) {
                read_restart_file(out + ".rst." + std::to_string(opt.rank), ctx, it_saved, n_saved);
            } else {   // the reference's own way: from the output files of the last save point
                if (n_cov) fatal("--restart from the output files with --covariates: only the .rst.<rank> state files carry the fixed effects");
                std::vector<int32_t> bs(opt.tasks), bl(opt.tasks);
                HB(hb_get_task_blocks(ctx, bs.data(), bl.data()));
                restart_from_outputs(out, ctx, Mtot, N, G, K, opt.rank * TL, TL, m_start, m_local, bl, opt.save, it_saved, n_saved);
            }
            {   // the per-process state files are written independently: all must stem from the same --save point
                const uint64_t v[2] = {it_saved, n_saved};
                HB(hb_comm_check_equal(ctx, v, 2, "--restart: the restart point (iteration, records) of the <out>.rst.<rank> files"));
            }
            it_first = it_saved + 1;
            if (root) {
                truncate_csv(out + ".csv", it_saved);
                csv.open_append(out + ".csv");
                if (truncate((out + ".bet").c_str(), 4 + (off_t)n_saved * (4 + (off_t)Mtot * 8)) != 0 ||
                    truncate((out + ".acu").c_str(), 4 + (off_t)n_saved * (4 + (off_t)Mtot * 8)) != 0 ||
                    truncate((out + ".cpn").c_str(), 4 + (off_t)n_saved * (4 + (off_t)Mtot * 4)) != 0)
                    fatal("--restart: cannot cut the .bet/.acu/.cpn files back to the restart point");
            }
            for (uint32_t t = 0; t < TL; t++) {
                const std::string mf = out + ".mus." + std::to_string(t_first + t);
                if (truncate(mf.c_str(), (off_t)n_saved * 12) != 0) fatal("--restart: cannot cut " + mf + " back to the restart point");
                mus[t].open_append(mf);
            }
            printf("INFO   : restarting after iteration %u (%u records in .bet/.cpn/.acu)\n", it_saved, n_saved);
        } else {
            for (uint32_t t = 0; t < TL; t++) mus[t].open(out + ".mus." + std::to_string(t_first + t));
        }

        std::vector<double> beta(m_local), acum(m_local), sigmaG(G), pi((size_t)G * K), mu(TL), bsq(G), eps(N);
        std::vector<int32_t> comp(m_local), cass((size_t)G * K), m0(G), perm;
        double tot_loop_ms = 0.0, tot_iter_ms = 0.0;
        for (uint32_t it = it_first; it < opt.chainLength; it++) {
            hb_brr_iter_out io;
            HB(hb_brr_iteration(ctx, nullptr, &io));
            tot_loop_ms += io.loop_ms; tot_iter_ms += io.iter_ms;
            double sigmaE = 0.0;
            HB(hb_brr_get_hyper(ctx, sigmaG.data(), pi.data(), &sigmaE, mu.data(), bsq.data(), cass.data(), m0.data()));
            double sG = 0.0; int m0s = 0;
            for (uint32_t g = 0; g < G; g++) { sG += sigmaG[g]; m0s += m0[g]; }
            if (opt.rank % 10 == 0)  // :2713
                printf("RESULT : it %4u, rank %4d: proc = %9.3f s, sync = %9.3f (%9.3f + %9.3f), n_sync = %8llu (%8llu + %8llu) (%7.3f / %7.3f), sigmaG = %15.10f, sigmaE = %15.10f, betasq = %15.10f, m0 = %10d\n",
                       it, (int)opt.rank, io.iter_ms * 1e-3, 0.0, 0.0, 0.0, (unsigned long long)io.n_sync, (unsigned long long)io.n_windows, (unsigned long long)io.n_sync,
                       0.0, 0.0, sG, sigmaE, bsq[0], m0s);  // :2714-2720
            if (it % opt.thin == 0) {
                if (root) {
                    char buff[65536];
                    int n = snprintf(buff, sizeof(buff), "%5d, %4d", (int)it, (int)G);  // :2742-2764
                    for (uint32_t g = 0; g < G; g++) n += snprintf(buff + n, sizeof(buff) - n, ", %20.15f", sigmaG[g]);
                    n += snprintf(buff + n, sizeof(buff) - n, ", %20.15f, %20.15f, %7d, %4d, %2d", sigmaE, sG / (sigmaE + sG), m0s, (int)G, (int)K);
                    for (size_t x = 0; x < (size_t)G * K; x++) n += snprintf(buff + n, sizeof(buff) - n, ", %20.15f", pi[x]);
                    n += snprintf(buff + n, sizeof(buff) - n, "\n");
                    csv.put(buff, (size_t)n);
                    fflush(csv.f);
                }
                HB(hb_brr_get_state(ctx, beta.data(), comp.data(), acum.data()));
                // records {u32 it; f64[Mtot]} / {u32 it; i32[Mtot]}; every process writes its slice at MrankS (:2768-2785)
                const size_t rec8 = 4 + (size_t)n_saved * (4 + (size_t)Mtot * 8), rec4 = 4 + (size_t)n_saved * (4 + (size_t)Mtot * 4);
                if (root) { bet.write_at(&it, 4, rec8); acu.write_at(&it, 4, rec8); cpn.write_at(&it, 4, rec4); }
                bet.write_at(beta.data(), (size_t)m_local * 8, rec8 + 4 + (size_t)m_start * 8);
                acu.write_at(acum.data(), (size_t)m_local * 8, rec8 + 4 + (size_t)m_start * 8);
                cpn.write_at(comp.data(), (size_t)m_local * 4, rec4 + 4 + (size_t)m_start * 4);
                for (uint32_t t = 0; t < TL; t++) { mus[t].put(&it, 1); mus[t].put(&mu[t], 1); fflush(mus[t].f); }  // :2788-2791
                n_saved++;
            }
            if (it > 0 && it % opt.save == 0) {  // :2808-2838, overwritten at every save
                std::vector<int32_t> bs(opt.tasks), bl(opt.tasks);
                HB(hb_get_task_blocks(ctx, bs.data(), bl.data()));
                for (uint32_t t = 0; t < TL; t++) {
                    HB(hb_brr_get_task_epsilon(ctx, t, eps.data()));
                    dump_file(out + ".eps." + std::to_string(t_first + t), it, N, eps.data());
                    perm.resize((size_t)bl[t_first + t]);
                    HB(hb_brr_get_task_perm(ctx, t, perm.data()));
                    dump_file(out + ".mrk." + std::to_string(t_first + t), it, (uint32_t)bl[t_first + t], perm.data());
                    size_t need = 0;                                        // .rng.<task>: the task's stream as text (:2805)
                    HB(hb_brr_get_task_rng(ctx, t, nullptr, 0, &need));
                    std::string rs(need, '\0');
                    HB(hb_brr_get_task_rng(ctx, t, &rs[0], need, &need));
                    std::ofstream rf(out + ".rng." + std::to_string(t_first + t), std::ios::out | std::ios::trunc | std::ios::binary);
                    rf << rs.c_str();
                }
                if (n_cov) {   // .gam.<rank> / .xiv.<rank>: u32 it; u32 len; f64 / i32 [len] (:2811-2831)
                    std::vector<double> gam(n_cov);
                    std::vector<int32_t> xiv(n_cov);
                    HB(hb_brr_get_gamma(ctx, gam.data(), xiv.data()));
                    dump_file(out + ".gam." + std::to_string(opt.rank), it, n_cov, gam.data());
                    dump_file(out + ".xiv." + std::to_string(opt.rank), it, n_cov, xiv.data());
                }
                HB(hb_brr_get_state(ctx, beta.data(), comp.data(), nullptr));
                // u32 Mtot; u32 it; data[Mtot] (:2818-2838)
                if (root) { xb.write_at(&it, 4, 4); xc.write_at(&it, 4, 4); }
                xb.write_at(beta.data(), (size_t)m_local * 8, 8 + (size_t)m_start * 8);
                xc.write_at(comp.data(), (size_t)m_local * 4, 8 + (size_t)m_start * 4);
                write_restart_file(out + ".rst." + std::to_string(opt.rank), ctx, it, n_saved);
                // tarball of the dump (:2851-2875): `tar -cf <dir>/tarballs/dump_<name>_<it>__<date>.tar -T <out>.lst` by rank 0 once all
                // processes have written; the reference runs tar whether or not <dir>/tarballs exists, here only where it does
                {
                    const uint64_t v[1] = {it};
                    HB(hb_comm_check_equal(ctx, v, 1, "the save point"));   // barrier (no-op on one GPU)
                    struct stat stb;
                    if (root && stat((opt.mcmcOutDir + "/tarballs").c_str(), &stb) == 0 && S_ISDIR(stb.st_mode)) {
                        const time_t now = time(nullptr);
                        const tm *ltm = localtime(&now);
                        char tar[1024];
                        snprintf(tar, sizeof(tar), "dump_%s_%05d__%4d-%02d-%02d_%02d-%02d-%02d.tar", opt.mcmcOutNam.c_str(), (int)it, 1900 + ltm->tm_year,
                                 1 + ltm->tm_mon, ltm->tm_mday, ltm->tm_hour, ltm->tm_min, ltm->tm_sec);
                        printf("INFO   : will create tarball %s in %s with file listed in %s.\n", tar, opt.mcmcOutDir.c_str(), (out + ".lst").c_str());
                        const std::string cmd = "tar -cf " + opt.mcmcOutDir + "/tarballs/" + tar + " -T " + out + ".lst 2>/dev/null";
                        if (system(cmd.c_str()) != 0) printf("WARNING: tar reported missing files (optional outputs such as .gam / .xiv are listed unconditionally, as in the reference)\n");
                    }
                }
            }
        }
        if (root)
        printf("INFO   : time to process the data: %.3f sec (marker loops %.3f sec: %.3f M marker updates/s)\n", tot_iter_ms * 1e-3,
               tot_loop_ms * 1e-3, (double)Mtot * (opt.chainLength - it_first) / (tot_loop_ms * 1e3));
        hb_destroy(ctx);
    } catch (const std::exception &e) {
        fatal(e.what());
    }
    return 0;
}
