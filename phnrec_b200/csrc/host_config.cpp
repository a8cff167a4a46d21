// host_config.cpp — reader for the <dir>/config INI dialect (configz.cpp:102-165) with the typed
// variable table of the reference orchestrator (srec.cpp:34-110): unknown variables and badly typed
// values are rejected with the reference's error classes (EN_UNKVAR / EN_BADVAL / EN_INVVAR).
#include "internal.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace phn {
namespace {
enum VarType { T_STR, T_INT, T_FLOAT, T_BOOL };
struct Var { const char *sec, *name; VarType type; const char *def; };

// Same variables, types and defaults as the reference accepts; the KWS / STK / PLP ones are
// accepted (so shipped config files parse unchanged) and otherwise ignored by this hot path.
const Var kVars[] = {
    {"source", "format", T_STR, "lin16"}, {"source", "sample_freq", T_INT, "8000"},
    {"source", "scale", T_FLOAT, "1.0f"}, {"source", "dc_shift", T_FLOAT, "0.0f"},
    {"source", "noise_level", T_FLOAT, "0.0f"},
    {"params", "kind", T_STR, "fbanks"}, {"params", "suffix", T_STR, "mel"},
    {"melbanks", "nbanks", T_INT, "15"}, {"melbanks", "nbanks_full", T_INT, "-1"},
    {"melbanks", "lower_freq", T_FLOAT, "0"}, {"melbanks", "higher_freq", T_FLOAT, "4000"},
    {"melbanks", "vector_size", T_INT, "200"}, {"melbanks", "vector_step", T_INT, "80"},
    {"melbanks", "preem_coef", T_FLOAT, "0.0"}, {"melbanks", "z_mean_source", T_BOOL, "false"},
    {"plp", "order", T_INT, "12"}, {"plp", "compress_fact", T_FLOAT, "0.3333333"},
    {"plp", "cep_lifter", T_FLOAT, "22"}, {"plp", "cep_scale", T_FLOAT, "10"}, {"plp", "add_c0", T_BOOL, "false"},
    {"onlinenorm", "estim_interval", T_INT, "0"}, {"onlinenorm", "signal_est_end", T_BOOL, "false"},
    {"onlinenorm", "file", T_STR, "none"}, {"onlinenorm", "mean_norm", T_BOOL, "false"},
    {"onlinenorm", "var_norm", T_BOOL, "false"}, {"onlinenorm", "scale_to_gvar", T_BOOL, "false"},
    {"offlinenorm", "sent_mean_norm", T_BOOL, "false"}, {"offlinenorm", "sent_var_norm", T_BOOL, "false"},
    {"offlinenorm", "sent_std_thr", T_FLOAT, "0.01"}, {"offlinenorm", "sent_max_norm", T_BOOL, "false"},
    {"offlinenorm", "sent_chmax_norm", T_BOOL, "false"},
    {"framenorm", "min_floor", T_FLOAT, "-9999.9"}, {"framenorm", "shift", T_FLOAT, "0"},
    {"posteriors", "system", T_STR, "1BT_DCT"}, {"posteriors", "length", T_INT, "31"},
    {"posteriors", "add_c0", T_BOOL, "true"}, {"posteriors", "hamming", T_BOOL, "false"},
    {"posteriors", "suffix", T_STR, "lop"}, {"posteriors", "bunch_size", T_STR, "1"},
    {"posteriors", "enabled", T_BOOL, "true"}, {"posteriors", "softening_func", T_STR, "none 0 0 0"},
    {"decoder", "type", T_STR, "stkint"}, {"decoder", "wpenalty", T_FLOAT, "-2.0"},
    {"decoder", "lm_scale", T_FLOAT, "1.0"}, {"decoder", "time_pruning", T_INT, "40"},
    {"decoder", "mode", T_STR, "decode"}, {"decoder", "softening_func", T_STR, "log 0 0 0"},
    {"decoder", "num_states_per_phn", T_INT, "1"},
    {"dirs", "tmp", T_STR, "$C/tmp"},
    {"models", "hmm_defs", T_STR, "$T/models"}, {"models", "nstates", T_INT, "3"},
    {"models", "gen_from_phn_list", T_BOOL, "false"},
    {"dicts", "phoneme_list", T_STR, ""}, {"dicts", "lexicon1", T_STR, ""}, {"dicts", "lexicon2", T_STR, ""},
    {"dicts", "lexicon1_save_bin", T_BOOL, "false"}, {"dicts", "lexicon2_save_bin", T_BOOL, "false"},
    {"dicts", "keyword_list", T_STR, "none"}, {"dicts", "charset", T_STR, "eastevrope"},
    {"networks", "default", T_STR, "$C/nets/network"}, {"networks", "gen_phn_loop", T_BOOL, "false"},
    {"networks", "gen_kws_net", T_BOOL, "false"}, {"networks", "omit_phn", T_STR, "oth"},
    {"labels", "suffix", T_STR, "rec"}, {"labels", "remove_path", T_BOOL, "true"},
    {"kws", "default_thr", T_FLOAT, "-10.0"}, {"kws", "thresholds_file", T_STR, "none"},
    {"gptransc", "rules", T_STR, "none"}, {"gptransc", "symbols", T_STR, "none"},
    {"gptransc", "max_variants", T_INT, "-1"}, {"gptransc", "scale_prob", T_BOOL, "false"},
    {"gptransc", "prob_thr", T_FLOAT, "-1.0"}, {"phntransc", "mode", T_STR, "lexgpt"},
};

// value syntax checks, as strict as configz.cpp:60-99 (sscanf-based) needs for the shipped files
int check_value(VarType t, const char *v)
{
    int iv;
    float fv;
    switch (t) {
        case T_INT: return sscanf(v, "%d", &iv) == 1 ? PHN_OK : PHN_ERR_CFG_BADVAL;
        case T_FLOAT: return sscanf(v, "%f", &fv) == 1 ? PHN_OK : PHN_ERR_CFG_BADVAL;
        case T_BOOL: return (!strcmp(v, "true") || !strcmp(v, "false")) ? PHN_OK : PHN_ERR_CFG_BADVAL;
        default: return PHN_OK;
    }
}
}  // namespace

int Config::load(const std::string &file, int *err_line)
{
    kv.clear();
    for (const Var &v : kVars) kv[std::string(v.sec) + "/" + v.name] = v.def;
    FILE *fp = fopen(file.c_str(), "rb");
    if (!fp) return PHN_ERR_CFG_FILE;
    char buf[1024];
    std::string section;
    int line = 1;
    while (fgets(buf, sizeof buf - 1, fp)) {
        size_t n = strcspn(buf, "\r\n");  // CR/LF tolerant: cut at the first of either
        buf[n] = 0;
        if (err_line) *err_line = line;
        if (n > 1 && buf[0] == '[') {
            section.assign(buf + 1, n - 2);  // drop the closing bracket
        } else if (n == 0 || buf[0] == '#') {
            // blank or comment
        } else {
            char *eq = strchr(buf, '=');
            // "var=value#comment": variable up to '=', value up to '#'; both must be non-empty
            if (!eq || eq == buf) { fclose(fp); return PHN_ERR_CFG_INVVAR; }
            *eq = 0;
            char *val = eq + 1;
            if (char *h = strchr(val, '#')) *h = 0;
            if (!*val) { fclose(fp); return PHN_ERR_CFG_INVVAR; }
            const Var *found = nullptr;
            for (const Var &v : kVars)
                if (section == v.sec && !strcmp(buf, v.name)) { found = &v; break; }
            if (!found) { fclose(fp); return PHN_ERR_CFG_UNKVAR; }
            int rc = check_value(found->type, val);
            if (rc != PHN_OK) { fclose(fp); return rc; }
            kv[section + "/" + buf] = val;
        }
        ++line;
    }
    fclose(fp);
    return PHN_OK;
}

const std::string &Config::str(const char *sec, const char *var) const
{
    static const std::string empty;
    auto it = kv.find(std::string(sec) + "/" + var);
    return it == kv.end() ? empty : it->second;
}
int Config::i(const char *sec, const char *var) const { return atoi(str(sec, var).c_str()); }
float Config::f(const char *sec, const char *var) const
{
    float v = 0.f;
    sscanf(str(sec, var).c_str(), "%f", &v);
    return v;
}
bool Config::b(const char *sec, const char *var) const { return str(sec, var) == "true"; }

}  // namespace phn
