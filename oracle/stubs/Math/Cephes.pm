# oracle/stubs/Math/Cephes.pm -- TEST INFRASTRUCTURE ONLY.
# Math::Cephes is not installed for this perl; the reference ships only a perl-5.10.1 PAR.
# This stub forwards bdtrc() through a tiny XS shim (oracle/xs/Cephes.xs) to the raw C symbol
# of the reference's own bundled binary (CEPHES_REAL_SO = oracle/_ref/Cephes.so) or, when
# CEPHES_REAL_SO points at libppp_oracle.so, to the C restatement (symbol cp_bdtrc).
package Math::Cephes;
use strict;
require Exporter;
require DynaLoader;
our @ISA       = qw(Exporter DynaLoader);
our @EXPORT_OK = qw(bdtrc);
our $VERSION   = '0.47';
my $so  = $ENV{CEPHES_SHIM_SO} or die "CEPHES_SHIM_SO unset";
my $lib = DynaLoader::dl_load_file( $so, 0 ) or die DynaLoader::dl_error();
my $sym = DynaLoader::dl_find_symbol( $lib, "boot_Math__Cephes" ) or die DynaLoader::dl_error();
my $xs  = DynaLoader::dl_install_xsub( "Math::Cephes::bootstrap_shim", $sym, $so );
&$xs("Math::Cephes");
_load( $ENV{CEPHES_REAL_SO}, $ENV{CEPHES_SYMBOL} || "bdtrc" ) or die "no bdtrc symbol";
1;
