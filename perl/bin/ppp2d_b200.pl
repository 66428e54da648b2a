#!/usr/bin/env perl
# perl/bin/ppp2d_b200.pl -- command-line drop-in for ProPhylo's bin/ppp2d.pl (two-dimensional PPP scan) on the B200 path.
#
# Same options, database layout and stdout as the reference's server mode (/root/reference/bin/ppp2d.pl:186-212, 228-364): the
# BLAST cache of gene <gi> is walked line by line; every line that brings a new taxon defines a depth-prefix query profile
# {taxa seen so far => 1, every other taxon of data/taxdis.txt => 0} with p = taxa / 1459 (:253-315); each prefix is scored
# against every gene of the genome, the best locus other than the gene itself is kept unless its printed score is 0.000
# (:322-352), and the rows are printed sorted by printed score (:354-364).  The reference forks `perl ppp.pl` once per depth
# (:317-320), which in turn forks its clients; here ALL depths x ALL genes are ONE GPU call (top-2 per depth) through
# PhyloProf::B200 (XS -> libppp_b200.so) and the native `.bla` reader.  gi -> taxon comes from the reference's own
# SeqToolBox::Taxonomy.  There is no CPU fallback.
#
#   perl ppp2d_b200.pl -d <db dir> -w <workspace> -g <gi> [-t taxon] [-s start depth] [-e end depth] [-j skip] [--dup]
#        [--taxdis FILE] [--device N | --devices 0-7]
# --taxdis defaults to <this script's dir>/../data/taxdis.txt, where the reference keeps it.  --threads/--serial/--keep/-c/-o/
# -r/-n/-p/--prob/-l are accepted and ignored as far as the reference's server mode ignores them too.
use strict;
use warnings;
use Getopt::Long;
use File::Spec;
use FindBin;
use lib "$FindBin::Bin/../lib";
use PhyloProf::B200;
use PhyloProf::B200::Host qw(log10_3f handle_negative_zero perl_num15);

my ( $help, $db_dir, $workspace, $taxon, $gi, $start, $end, $skip, $duplicate, $taxdis_file, $device, $devices ) = ( 0, '', '', 0, '', 5, undef, 1, 0, '', 0, '' );
my %ignored;
GetOptions(
    'help|h' => \$help, 'db|d=s' => \$db_dir, 'workspace|w=s' => \$workspace, 'taxon|t=i' => \$taxon, 'gi|g=i' => \$gi,
    'start-depth|s=i' => \$start, 'end-depth|e=i' => \$end, 'skip|j=i' => \$skip, 'dup' => \$duplicate, 'taxdis=s' => \$taxdis_file,
    'device=i' => \$device, 'devices=s' => \$devices,
    'profile|p=s' => \$ignored{p}, 'level|l=s' => \$ignored{l}, 'project|r=i' => \$ignored{r}, 'nodes|n=i' => \$ignored{n}, 'client' => \$ignored{client},
    'copy|c' => \$ignored{c}, 'output|o=s' => \$ignored{o}, 'prob=s' => \$ignored{prob}, 'serial' => \$ignored{serial}, 'threads=i' => \$ignored{threads},
    'keep|k' => \$ignored{k},
) or die "usage: ppp2d_b200.pl -d <db> -w <workspace> -g <gi> [-s start] [-e end] [-j skip]\n";
die "usage: ppp2d_b200.pl -d <db> -w <workspace> -g <gi> [-s start] [-e end] [-j skip]\n" if $help || !$gi;
die "ppp2d_b200.pl: --client is the reference's worker mode; this driver has no workers\n" if $ignored{client};
my $total_taxa_count = 1459;    # bin/ppp2d.pl:175: hard-coded in the reference

eval { require SeqToolBox::Taxonomy; 1 } or die "ppp2d_b200.pl: needs SeqToolBox::Taxonomy (the reference's lib/) on \@INC: $@";
my $input_taxonomy = SeqToolBox::Taxonomy->new()->get_taxon($gi);
$taxon = $input_taxonomy unless $taxon;
die "Can't file taxon for $gi\n" unless $input_taxonomy;
my $blast_result_file = File::Spec->catfile( $db_dir, "blast", $input_taxonomy, "$gi.bla" );
die "Can't locate $blast_result_file for $gi\n" unless -s $blast_result_file;

$taxdis_file ||= File::Spec->catfile( $FindBin::Bin, "..", "data", "taxdis.txt" );
open( my $td, "<", $taxdis_file ) or die "Can't open $taxdis_file\n";
my @taxdis = map { chomp; $_ } <$td>;
close($td);

# ---- the depth-prefix queries (:253-315)
open( my $bf, "<", $blast_result_file ) or die "Can't open $blast_result_file\n";
my ( @queries, @seen, %seen );
my $line_count = 0;
while ( my $line = <$bf> ) {
    chomp $line;
    ++$line_count;
    my @f = split( /\t/, $line );
    no warnings 'uninitialized';
    if ( $f[0] eq $f[1] ) {    # the hit to itself only marks its taxon
        push @seen, $f[3] unless $seen{ $f[3] }++;
        next;
    }
    next if $seen{ $f[3] }++;
    push @seen, $f[3];
    next if $line_count % $skip;
    my $tc = scalar @seen;
    last if $end && $tc > $end;
    push @queries, [ $line_count, [@seen], perl_num15( $tc / $total_taxa_count ) ] if $tc >= $start;    # --prob goes through a command line: %.15g
}
close($bf);
print "#Subject_best_depth\tLocus\t\#Yes\t#Total\tDepth\tProb\tScore\tDesc\n";
exit(0) unless @queries;

# ---- the genome's gene caches (what the forked ppp.pl lists, bin/ppp.pl:955-1017)
my $blast_dir = File::Spec->catdir( $db_dir, "blast", $taxon );
opendir( my $dh, $blast_dir ) or die "Can't find blast result. Did you setup the database properly?\n";
my ( @loci, @paths );
foreach my $file ( sort readdir($dh) ) {
    next unless $file =~ /(\S+)\.bla$/;
    push @loci,  $1;
    push @paths, File::Spec->catfile( $blast_dir, $file );
}
closedir($dh);
die "Can't find blast result. Did you setup the database properly?\n" unless @paths;

# ---- one GPU call: every prefix x every gene, best 2 per prefix (the best one may be the gene itself)
my %all = map { $_ => 1 } @taxdis, @seen;
delete $all{''};
my @universe = sort keys %all;
my @profiles;
foreach my $q (@queries) {
    my %h = map { $_ => 0 } grep { $_ ne '' } @taxdis;
    $h{$_} = 1 for grep { $_ ne '' } @{ $q->[1] };
    push @profiles, \%h;
}
my $gpu = $devices ne '' ? PhyloProf::B200->new( -devices => $devices ) : PhyloProf::B200->new( -device => $device );
my $k = @paths < 2 ? scalar(@paths) : 2;
my $hits = $gpu->score_bla_files( \@universe, \@profiles, \@paths, -probs => [ map { $_->[2] } @queries ], -filter => 'topk', -k => $k, -dup => $duplicate );
my @best;    # per query: its rows in ppp.pl's order (ascending raw score)
push @{ $best[ $_->[0] ] }, $_ for @$hits;

# ---- descriptions are optional, as in ppp.pl (:468-506); an empty Desc is a trailing empty field, which the reference's
# split-and-join of the row drops (:333-345)
my $sth;
my $desc_db = File::Spec->catfile( $db_dir, "desc", "gi_desc.sqlite3" );
if ( -s $desc_db && eval { require DBI; 1 } ) {
    my $dbh = eval { DBI->connect( "dbi:SQLite:dbname=$desc_db", "", "", { RaiseError => 1, AutoCommit => 0 } ) };
    $sth = eval { $dbh->prepare('select desc from gi_desc where gi =?') } if $dbh;
}
my @rows;
for my $qi ( 0 .. $#queries ) {
    my ($h) = grep { $loci[ $_->[1] ] ne $gi } sort { perl_num15( $a->[2] ) <=> perl_num15( $b->[2] ) || $loci[ $a->[1] ] cmp $loci[ $b->[1] ] } @{ $best[$qi] || [] };
    next unless $h;
    my $score = handle_negative_zero( log10_3f( perl_num15( $h->[2] ) ) );
    next if $score == 0.00;    # the reference compares the PRINTED score (:340)
    my @f = ( $loci[ $h->[1] ], $h->[3], $h->[4], $h->[5], sprintf( "%.3f", perl_num15( $queries[$qi][2] ) ), $score );
    if ($sth) { $sth->execute( $f[0] ); my @r = $sth->fetchrow_array(); push @f, $r[0] if $r[0]; }
    push @rows, [ $queries[$qi][0], \@f ];
}
foreach my $r ( sort { $a->[1][5] <=> $b->[1][5] || $a->[0] <=> $b->[0] } @rows ) {
    print $r->[0], "\t", join( "\t", @{ $r->[1] } ), "\n";
}
exit(0);
